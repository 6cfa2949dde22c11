"""Host-side operators over the C ABI (include/mmrec_b200.h): the three op families of MMRec's hot path.

Each function here replaces a PyTorch library call of the reference (cited per function, paths relative
to /root/reference) with a call into libmmrec_b200.so on the current CUDA stream.  PyTorch is used for
device memory, streams and autograd bookkeeping only.  There is no CPU path: tensors must live on a
sm_100 device, otherwise `MMRecError` is raised.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import MMRecError, check

_ws_cache: dict = {}
import os as _os
SEG = int(_os.environ.get("MMREC_SPMM_SEG", "512"))        # non-zeros per SpMM task (rows longer than this are split)
LIGHT_MAX = int(_os.environ.get("MMREC_SPMM_LIGHT", "32"))  # tasks longer than this are run by a whole CTA
# All SpMMs of an inference-time propagation in one cooperative launch (mmrec_spmm_chain_f32).  Measured on B200 at the baby
# graph (profiles/r02_notes.md): 82.5 us against 75.8 us for one launch per SpMM replayed from a CUDA graph -- a grid-wide
# barrier costs more than the launch boundary it replaces -- so it is opt-in.
CHAIN = _os.environ.get("MMREC_SPMM_CHAIN", "0") == "1"


def launch_count() -> int:
    """Kernels of this library launched by this process so far: every launch site in the C ABI counts itself
    (`mmrec_launch_count`); bench.py's `gpu_launches` is the difference over its timed region (for a section replayed
    from a CUDA graph: the launches recorded while capturing it, times the replays)."""
    return int(_lib.load().mmrec_launch_count())


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise MMRecError("mmrec_b200 ops run on CUDA tensors only (no CPU fallback)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise MMRecError(f"expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ws(name: str, nbytes: int, device) -> torch.Tensor:
    """Scratch for one op family, per device AND per stream (two streams never share a buffer).  A buffer that turned
    out too small is kept alive next to its replacement: a CUDA graph captured earlier has its address baked in."""
    key = (name, device.index, torch.cuda.current_stream(device).cuda_stream)
    bufs = _ws_cache.setdefault(key, [])
    if not bufs or bufs[-1].numel() < nbytes:
        grow = int(bufs[-1].numel() * 1.5) if bufs else 0
        bufs.append(torch.empty(max(int(nbytes), grow, 256), dtype=torch.uint8, device=device))
    return bufs[-1]


# ------------------------------------------------------------------------------------------------
# K1c: CSR container
# ------------------------------------------------------------------------------------------------
class CSR:
    """Row-sorted int32 CSR of a sparse matrix on the device, with the SpMM work plan.

    Stands in for the reference's `torch.sparse` COO tensors (`norm_adj`, `masked_adj`, `mm_adj`, `R`;
    SURVEY.md 8a a1/a2/a8).  Built once per graph (per epoch for FREEDOM's pruned graph) instead of the
    coalesce + COO->CSR conversion ATen performs inside every `torch.sparse.mm` call.
    """

    def __init__(self, n_rows, n_cols, rowptr, colidx, vals, nnz, symmetric=False, seg=None, light_max=None):
        self.n_rows, self.n_cols, self.nnz = int(n_rows), int(n_cols), int(nnz)
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.symmetric = symmetric
        self.seg = SEG if seg is None else seg
        self.light_max = LIGHT_MAX if light_max is None else light_max
        self._t: Optional["CSR"] = None
        self._partial = {}
        self._plan()

    # -- construction ------------------------------------------------------------------------------
    @staticmethod
    def from_coo(row: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], n_rows: int, n_cols: int,
                 sum_duplicates: bool = True, symmetric: bool = False, seg=None, light_max=None) -> "CSR":
        _lib.require_device()
        _need_cuda(row, col, val)
        lib = _lib.load()
        row, col = row.to(torch.int64).contiguous(), col.to(torch.int64).contiguous()
        val = None if val is None else _f32c(val)
        nnz = row.numel()
        dev = row.device
        rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
        colidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        vals = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        nnz_out = torch.zeros(1, dtype=torch.int64, device=dev)
        nbytes = lib.mmrec_csr_from_coo_workspace_bytes(nnz, n_rows)
        ws = _ws("csr", nbytes, dev)
        check(lib.mmrec_csr_from_coo(nnz, _ptr(row), _ptr(col), _ptr(val), n_rows, n_cols, int(sum_duplicates),
                                     _ptr(rowptr), _ptr(colidx), _ptr(vals), _ptr(nnz_out), _ptr(ws), ws.numel(),
                                     _stream()), "mmrec_csr_from_coo")
        n = int(nnz_out.item())
        return CSR(n_rows, n_cols, rowptr, colidx[:max(n, 1)], vals[:max(n, 1)], n, symmetric, seg, light_max)

    @staticmethod
    def from_torch_sparse(t: torch.Tensor, symmetric: bool = False) -> "CSR":
        """From an (un-coalesced) torch COO tensor as the reference builds them."""
        idx, val = t._indices(), t._values()
        return CSR.from_coo(idx[0], idx[1], val.to(torch.float32), t.shape[0], t.shape[1], True, symmetric)

    def _plan(self):
        lib = _lib.load()
        dev = self.rowptr.device
        max_tasks = self.n_rows + self.nnz // self.seg + 1
        max_split = self.nnz // self.seg + 1
        tasks = torch.empty(4 * max_tasks, dtype=torch.int32, device=dev)
        split = torch.empty(4 * max_split, dtype=torch.int32, device=dev)
        counts = torch.zeros(8, dtype=torch.int64, device=dev)
        ws = _ws("plan", lib.mmrec_spmm_plan_workspace_bytes(self.n_rows, max_tasks), dev)
        check(lib.mmrec_spmm_plan(self.n_rows, _ptr(self.rowptr), self.seg, self.light_max, max_tasks, _ptr(tasks), _ptr(split),
                                  _ptr(counts), _ptr(ws), ws.numel(), _stream()), "mmrec_spmm_plan")
        c = counts.tolist()
        self.n_tasks, self.n_split, self.n_slots, self.longest_row = int(c[0]), int(c[1]), int(c[2]), int(c[3])
        self.n_cta_tasks = int(c[4])
        self.tasks = tasks[:4 * max(self.n_tasks, 1)]
        self.split_rows = split[:4 * max(self.n_split, 1)]
        self.counters = torch.zeros(max(self.n_split, 1), dtype=torch.int32, device=dev)

    def partial(self, d: int) -> torch.Tensor:
        t = self._partial.get(d)
        if t is None:
            t = torch.empty(max(self.n_slots, 1) * d, dtype=torch.float32, device=self.rowptr.device)
            self._partial[d] = t
        return t

    # -- views ---------------------------------------------------------------------------------------
    def coo(self):
        """(row int64, col int64, val) of the stored entries."""
        counts = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
        row = torch.repeat_interleave(torch.arange(self.n_rows, device=self.rowptr.device), counts)
        return row, self.colidx[:self.nnz].to(torch.int64), self.vals[:self.nnz]

    def t(self) -> "CSR":
        """Transposed matrix (needed by the backward of directed graphs: mm_adj, R)."""
        if self.symmetric:
            return self
        if self._t is None:
            r, c, v = self.coo()
            self._t = CSR.from_coo(c, r, v, self.n_cols, self.n_rows, False, False, self.seg, self.light_max)
            self._t._t = self
        return self._t

    def to_dense(self) -> torch.Tensor:
        r, c, v = self.coo()
        out = torch.zeros(self.n_rows, self.n_cols, dtype=torch.float32, device=v.device)
        out.index_put_((r, c), v, accumulate=True)
        return out

    def algorithmic_bytes(self, d: int) -> int:
        """4(n_rows+1) + 8 nnz + 4 n_cols d + 4 n_rows d."""
        return 4 * (self.n_rows + 1) + 8 * self.nnz + 4 * self.n_cols * d + 4 * self.n_rows * d


# ------------------------------------------------------------------------------------------------
# K1: SpMM
# ------------------------------------------------------------------------------------------------
class PanelCSR:
    """A sparse matrix cut into column panels (each an ordinary `CSR` over all rows and the panel's columns only) for graphs
    whose dense operand does not fit the L2 (BASELINE config 5: 2M x 128 fp32 user table = 1 GB).  `spmm_raw` multiplies
    the panels one after the other, accumulating in Y, so that the rows of X a panel gathers -- `panel_bytes` of them -- are
    L2-resident: every row of X comes from HBM once per product instead of once per non-zero.  Same interface as `CSR` as far
    as the propagation needs it (`n_rows`, `n_cols`, `nnz`, `t()`, `algorithmic_bytes`)."""

    def __init__(self, n_rows, n_cols, panels, bounds, symmetric=False):
        self.n_rows, self.n_cols, self.panels, self.bounds, self.symmetric = n_rows, n_cols, panels, bounds, symmetric
        self.nnz = sum(p.nnz for p in panels)
        self.n_tasks = min(p.n_tasks for p in panels) if panels else 0
        self._t = None

    @staticmethod
    def from_coo(row, col, val, n_rows, n_cols, d, panel_bytes=48 << 20, sum_duplicates=True, symmetric=False) -> "PanelCSR":
        cols_per_panel = max(1024, panel_bytes // (4 * d))
        n_panels = max(1, -(-n_cols // cols_per_panel))
        cols_per_panel = -(-n_cols // n_panels)
        panels, bounds = [], []
        for p in range(n_panels):
            lo, hi = p * cols_per_panel, min(n_cols, (p + 1) * cols_per_panel)
            m = (col >= lo) & (col < hi)
            panels.append(CSR.from_coo(row[m], col[m], None if val is None else val[m], n_rows, n_cols, sum_duplicates, False))
            bounds.append((lo, hi))
        out = PanelCSR(n_rows, n_cols, panels, bounds, symmetric)
        out._coo = (row, col, val, d, panel_bytes, sum_duplicates)
        return out

    def t(self):
        if self.symmetric:
            return self
        if self._t is None:
            row, col, val, d, pb, sd = self._coo
            self._t = PanelCSR.from_coo(col, row, val, self.n_cols, self.n_rows, d, pb, sd, False)
            self._t._t = self
        return self._t

    def algorithmic_bytes(self, d: int) -> int:
        return 4 * (self.n_rows + 1) + 8 * self.nnz + 4 * self.n_cols * d + 4 * self.n_rows * d


def spmm_raw(A, X: torch.Tensor, Y: Optional[torch.Tensor] = None, acc_in: Optional[torch.Tensor] = None,
             acc_out: Optional[torch.Tensor] = None, acc_div: float = 1.0, gate_ref: Optional[torch.Tensor] = None,
             use_plan: bool = True, y_accumulate: bool = False):
    """y = A X with the fused epilogue of include/mmrec_b200.h (no autograd).  Replaces `torch.sparse.mm`
    (`src/models/freedom.py:167,172`) plus the stack/mean (`:175-176`) and `+ h` (`:178`) that follow."""
    if isinstance(A, PanelCSR):
        if gate_ref is not None:
            raise MMRecError("spmm: the cosine gate needs the whole row sum: not available on a PanelCSR")
        last = len(A.panels) - 1
        for i, P in enumerate(A.panels):                              # Y accumulates over the panels; the running sum takes every
            run_in = None if acc_out is None else (acc_in if i == 0 else acc_out)   # panel's share, the division comes with the last one
            spmm_raw(P, X, Y=Y, acc_in=run_in, acc_out=acc_out, acc_div=acc_div if i == last else 1.0, use_plan=use_plan,
                     y_accumulate=Y is not None and (i > 0 or y_accumulate))
        return
    _need_cuda(X, Y, acc_in, acc_out, gate_ref)
    lib = _lib.load()
    if X.dim() != 2 or X.shape[0] != A.n_cols:
        raise MMRecError(f"spmm: X is {tuple(X.shape)}, matrix has {A.n_cols} columns")
    X = _f32c(X)
    d = X.shape[1]
    for name, t in (("Y", Y), ("acc_in", acc_in), ("acc_out", acc_out), ("gate_ref", gate_ref)):
        if t is not None and (t.shape != (A.n_rows, d) or not t.is_contiguous() or t.dtype != torch.float32):
            raise MMRecError(f"spmm: {name} must be contiguous float32 [{A.n_rows}, {d}]")
    if Y is None and acc_out is None:
        raise MMRecError("spmm: nothing to write")
    plan = use_plan and A.n_tasks > 0
    if y_accumulate:
        if gate_ref is not None or Y is None:
            raise MMRecError("spmm: y_accumulate needs Y and no gate")
        check(lib.mmrec_spmm_acc_f32(A.n_rows, A.n_cols, d, _ptr(A.rowptr), _ptr(A.colidx), _ptr(A.vals),
                                     _ptr(A.tasks) if plan else None, A.n_tasks if plan else 0, A.n_cta_tasks if plan else 0,
                                     _ptr(A.split_rows) if plan else None, _ptr(A.counters) if plan else None,
                                     _ptr(A.partial(d)) if plan else None,
                                     _ptr(X), X.stride(0), _ptr(Y), d, _ptr(acc_in), _ptr(acc_out), d, float(acc_div), 1, _stream()),
              "mmrec_spmm_acc_f32")
        return
    check(lib.mmrec_spmm_f32(A.n_rows, A.n_cols, d, _ptr(A.rowptr), _ptr(A.colidx), _ptr(A.vals),
                             _ptr(A.tasks) if plan else None, A.n_tasks if plan else 0, A.n_cta_tasks if plan else 0,
                             _ptr(A.split_rows) if plan else None, _ptr(A.counters) if plan else None,
                             _ptr(A.partial(d)) if plan else None,
                             _ptr(X), X.stride(0), _ptr(Y), d, _ptr(acc_in), _ptr(acc_out), d, float(acc_div),
                             _ptr(gate_ref), d, _stream()), "mmrec_spmm_f32")


class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, A: CSR, base):
        ctx.A = A
        ctx.has_base = base is not None
        out = torch.empty(A.n_rows, X.shape[1], dtype=torch.float32, device=X.device)
        if base is None:
            spmm_raw(A, X, Y=out)
        else:
            spmm_raw(A, X, acc_in=_f32c(base), acc_out=out)
        return out

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        At = ctx.A.t()
        gx = torch.empty(At.n_rows, g.shape[1], dtype=torch.float32, device=g.device)
        spmm_raw(At, g, Y=gx)                       # dX = A^T dY  (autograd of torch.sparse.mm)
        return gx, None, (g if ctx.has_base else None)


def spmm(A: CSR, X: torch.Tensor, base: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`base + A @ X` (base optional), differentiable w.r.t. X and base."""
    return _SpmmFn.apply(X, A, base)


def _chain_step(A: CSR, X, Y=None, acc_in=None, acc_out=None, acc_div=1.0, post=None, post_row0=0, sync_before=False):
    d = X.shape[1]
    st = _lib.SpmmStep()
    st.n_rows, st.n_cols = A.n_rows, A.n_cols
    st.rowptr, st.colidx, st.vals = _ptr(A.rowptr), _ptr(A.colidx), _ptr(A.vals)
    st.tasks, st.n_tasks, st.n_cta_tasks = _ptr(A.tasks), A.n_tasks, A.n_cta_tasks
    st.split_rows, st.counters, st.partial = _ptr(A.split_rows), _ptr(A.counters), _ptr(A.partial(d))
    st.X, st.ldx = _ptr(X), X.stride(0)
    st.Y, st.ldy = _ptr(Y), d
    st.acc_in, st.acc_out, st.ldacc, st.acc_div = _ptr(acc_in), _ptr(acc_out), d, float(acc_div)
    st.post, st.ldpost, st.post_row0 = _ptr(post), d, int(post_row0)
    st.sync_before = int(bool(sync_before))
    return st


def propagate_mean_fused(A: CSR, ego: torch.Tensor, n_layers: int, post_csr: Optional[CSR] = None, post_x: Optional[torch.Tensor] = None,
                         post_layers: int = 1, post_row0: int = 0) -> torch.Tensor:
    """Inference form of `propagate_mean` (+ FREEDOM / BM3's item-item term) as ONE persistent cooperative launch
    (`mmrec_spmm_chain_f32`): `mean(E_0 .. E_L)`, and if `post_csr` is given `out[post_row0:] += post_csr^post_layers @ post_x`
    (`src/models/freedom.py:164-178`: `h = mm_adj @ ... @ item_emb`, `i_g + h`).  No autograd.  Falls back to one launch per
    SpMM when the chained kernel does not take the shape."""
    import ctypes
    _need_cuda(ego, post_x)
    ego = _f32c(ego)
    d = ego.shape[1]
    if n_layers < 1 or A.n_tasks == 0 or (post_csr is not None and post_csr.n_tasks == 0):
        return _propagate_mean_post_unfused(A, ego, n_layers, post_csr, post_x, post_layers, post_row0)
    steps, keep = [], []
    h = None
    if post_csr is not None:
        h = _f32c(post_x)
        for i in range(post_layers):                                 # h = mm_adj @ h, the first one reads the parameters only
            y = torch.empty(post_csr.n_rows, d, dtype=torch.float32, device=ego.device)
            steps.append(_chain_step(post_csr, h, Y=y, sync_before=i > 0))
            keep.append(y)
            h = y
    acc = torch.empty_like(ego)
    x = ego
    for l in range(1, n_layers + 1):
        last = l == n_layers
        y = None if last else torch.empty_like(ego)
        steps.append(_chain_step(A, x, Y=y, acc_in=ego if l == 1 else acc, acc_out=acc, acc_div=float(n_layers + 1) if last else 1.0,
                                 post=h if last else None, post_row0=post_row0,
                                 sync_before=(l > 1) or (last and h is not None)))
        keep.append(y)
        x = y
    if len(steps) > 8:
        return _propagate_mean_post_unfused(A, ego, n_layers, post_csr, post_x, post_layers, post_row0)
    arr = (_lib.SpmmStep * len(steps))(*steps)
    rc = _lib.load().mmrec_spmm_chain_f32(d, len(steps), ctypes.cast(arr, ctypes.c_void_p), _stream())
    if rc == -4:                                                     # MMREC_EUNSUPPORTED: shape without a chained kernel
        return _propagate_mean_post_unfused(A, ego, n_layers, post_csr, post_x, post_layers, post_row0)
    check(rc, "mmrec_spmm_chain_f32")
    return acc


def _propagate_mean_post_unfused(A, ego, n_layers, post_csr, post_x, post_layers, post_row0):
    out = _PropagateMeanFn.apply(ego.detach(), A, n_layers)
    if post_csr is not None:
        h = _f32c(post_x).detach()
        for _ in range(post_layers - 1):
            y = torch.empty(post_csr.n_rows, h.shape[1], dtype=torch.float32, device=h.device)
            spmm_raw(post_csr, h, Y=y)
            h = y
        tail = out[post_row0:]
        spmm_raw(post_csr, h, acc_in=tail, acc_out=tail)
    return out


class _PropagateMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ego, A: CSR, n_layers: int):
        ctx.A, ctx.L = A, n_layers
        ego = _f32c(ego)
        if n_layers == 0:
            return ego.clone()
        acc = torch.empty_like(ego)
        x = ego
        for l in range(1, n_layers + 1):
            last = l == n_layers
            y = None if last else torch.empty_like(ego)
            spmm_raw(A, x, Y=y, acc_in=ego if l == 1 else acc, acc_out=acc, acc_div=float(n_layers + 1) if last else 1.0)
            x = y
        return acc

    @staticmethod
    def backward(ctx, g):
        L = ctx.L
        gm = _f32c(g) / float(L + 1)                # d mean / d E_l, the same for every layer
        if L == 0:
            return g, None, None
        At = ctx.A.t()
        cur = gm
        for _ in range(L):                          # g_l = gm + A^T g_{l+1}
            nxt = torch.empty_like(gm)
            spmm_raw(At, cur, acc_in=gm, acc_out=nxt)
            cur = nxt
        return cur, None, None


def propagate_mean(A: CSR, ego: torch.Tensor, n_layers: int) -> torch.Tensor:
    """mean(E_0 .. E_L), E_{l+1} = A E_l -- the LightGCN propagation every graph model repeats
    (`src/models/freedom.py:169-176`, `bm3.py:86-92`, `lightgcn.py:116-123`, `mgcn.py:159-166`), with the
    running sum and the final division fused into the SpMM epilogue (no stack, no extra passes)."""
    if CHAIN and n_layers >= 1 and isinstance(A, CSR) and not (torch.is_grad_enabled() and ego.requires_grad):
        return propagate_mean_fused(A, ego, n_layers)               # inference: all layers in one cooperative launch
    return _PropagateMeanFn.apply(ego, A, n_layers)


def propagate_layergcn(A: CSR, ego: torch.Tensor, n_layers: int) -> torch.Tensor:
    """Inference form of `src/models/layergcn.py:125-138`: E_{l+1} = cos(A E_l, E_0) * A E_l, sum over layers
    1..L, gate and running sum fused into the SpMM epilogue.  (Training goes through `spmm` + torch ops so
    that autograd sees the cosine gate.)"""
    ego = _f32c(ego)
    if n_layers <= 0:
        return torch.zeros_like(ego)                 # the sum over layers 1..L of nothing
    acc = torch.empty_like(ego)
    x = ego
    for l in range(1, n_layers + 1):
        y = torch.empty_like(ego)
        spmm_raw(A, x, Y=y, acc_in=None if l == 1 else acc, acc_out=acc, gate_ref=ego)
        x = y
    return acc


# ------------------------------------------------------------------------------------------------
# K2: modality projection
# ------------------------------------------------------------------------------------------------
def project_raw(table, weight, bias, idx=None, l2_normalize=False) -> torch.Tensor:
    _need_cuda(table, weight, bias, idx)
    lib = _lib.load()
    table, weight = _f32c(table), _f32c(weight)
    bias = None if bias is None else _f32c(bias)
    if idx is not None:
        idx = idx.to(torch.int64).contiguous()
    n_out = table.shape[0] if idx is None else idx.numel()
    d, F = weight.shape
    if table.shape[1] != F:
        raise MMRecError(f"project: table has {table.shape[1]} features, weight expects {F}")
    out = torch.empty(n_out, d, dtype=torch.float32, device=table.device)
    ws = _ws("project", lib.mmrec_project_workspace_bytes(n_out, F, d), table.device)
    check(lib.mmrec_project_f32(n_out, _ptr(idx), _ptr(table), table.shape[0], F, _ptr(weight), _ptr(bias), d,
                                int(l2_normalize), _ptr(out), d, _ptr(ws), ws.numel(), _stream()), "mmrec_project_f32")
    return out


def set_project_path(tensor_core: bool):
    """True (default): tcgen05 3xTF32 projection kernel; False: exact fp32 CUDA-core kernel."""
    _lib.load().mmrec_project_set_path(int(bool(tensor_core)))


# -- f1: backward of the projection and the optimiser step (csrc/train.cu) -------------------------------------------
def index_sum_rows(g: torch.Tensor, idx: torch.Tensor, n_rows: int) -> torch.Tensor:
    """G[i] = sum_{j: idx[j] = i} g[j] in ascending j (`mmrec_index_sum_rows_f32`): `zeros.index_add_(0, idx, g)` made
    bit-reproducible.  The table gradient of a gathered projection is `G @ W` (linearity), so the scatter is d wide."""
    _need_cuda(g, idx)
    g = _f32c(g)
    idx = idx.to(torch.int64).contiguous()
    G = torch.empty(n_rows, g.shape[1], dtype=torch.float32, device=g.device)
    check(_lib.load().mmrec_index_sum_rows_f32(idx.numel(), _ptr(idx), _ptr(g), g.stride(0), g.shape[1], n_rows, _ptr(G), G.stride(0),
                                               _stream()), "mmrec_index_sum_rows_f32")
    return G


def linear_wgrad(g: torch.Tensor, table: torch.Tensor, idx: Optional[torch.Tensor] = None, want_bias: bool = True):
    """(dW [d, F], db [d] | None) of `y = table[idx] @ W^T + b` for the upstream gradient g [n, d]: `g.t().mm(x)`, `g.sum(0)`
    (autograd of `nn.Linear`, src/models/freedom.py:205-209) through `mmrec_linear_wgrad_f32`."""
    _need_cuda(g, table, idx)
    lib = _lib.load()
    g, table = _f32c(g), _f32c(table)
    if idx is not None:
        idx = idx.to(torch.int64).contiguous()
    n, d = g.shape
    F = table.shape[1]
    dW = torch.empty(d, F, dtype=torch.float32, device=g.device)
    db = torch.empty(d, dtype=torch.float32, device=g.device) if want_bias else None
    ws = _ws("wgrad", lib.mmrec_linear_wgrad_workspace_bytes(n, F, d), g.device)
    check(lib.mmrec_linear_wgrad_f32(n, _ptr(idx), _ptr(g), g.stride(0), d, _ptr(table), table.shape[0], F, _ptr(dW), _ptr(db),
                                     _ptr(ws), ws.numel(), _stream()), "mmrec_linear_wgrad_f32")
    return dW, db


def linear_dgrad(G: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """`G @ W` as a fresh [n_rows, F] tensor (`mmrec_linear_dgrad_f32`): the dense table gradient autograd expects."""
    _need_cuda(G, weight)
    G, weight = _f32c(G), _f32c(weight)
    out = torch.empty(G.shape[0], weight.shape[1], dtype=torch.float32, device=G.device)
    check(_lib.load().mmrec_linear_dgrad_f32(G.shape[0], _ptr(G), G.stride(0), G.shape[1], _ptr(weight), weight.shape[1], _ptr(out),
                                             _stream()), "mmrec_linear_dgrad_f32")
    return out


def linear_dgrad_adam(G, weight, param, exp_avg, exp_avg_sq, beta1, beta2, eps, weight_decay, step_size, bc2_sqrt):
    """One Adam step of `param` [n_rows, F] with the gradient `G @ W` computed inside the kernel, never stored
    (`mmrec_linear_dgrad_adam_f32`; torch.optim.Adam.step of src/common/trainer.py:189 fused with the projection backward)."""
    _need_cuda(G, weight, param, exp_avg, exp_avg_sq)
    G, weight = _f32c(G), _f32c(weight)
    for t in (param, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.shape != (G.shape[0], weight.shape[1]):
            raise MMRecError("linear_dgrad_adam: param / exp_avg / exp_avg_sq must be contiguous float32 [n_rows, F]")
    check(_lib.load().mmrec_linear_dgrad_adam_f32(G.shape[0], _ptr(G), G.stride(0), G.shape[1], _ptr(weight), weight.shape[1], _ptr(param),
                                                  _ptr(exp_avg), _ptr(exp_avg_sq), float(beta1), float(beta2), float(eps),
                                                  float(weight_decay), float(step_size), float(bc2_sqrt), _stream()),
          "mmrec_linear_dgrad_adam_f32")


def adam_step(entries, beta1, beta2, eps, weight_decay):
    """`entries`: (param, grad, exp_avg, exp_avg_sq, step_size, bc2_sqrt) per tensor; all updated by `mmrec_adam_f32`."""
    import ctypes
    if not entries:
        return
    arr = (_lib.AdamTensor * len(entries))()
    for a, (p, g, m, v, step_size, bc2_sqrt) in zip(arr, entries):
        _need_cuda(p, g, m, v)
        for t in (p, g, m, v):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != p.numel():
                raise MMRecError("adam_step: contiguous float32 tensors of one size per entry")
        a.param, a.grad, a.exp_avg, a.exp_avg_sq, a.n = _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel()
        a.step_size, a.bc2_sqrt = float(step_size), float(bc2_sqrt)
    check(_lib.load().mmrec_adam_f32(len(entries), ctypes.cast(arr, ctypes.c_void_p), float(beta1), float(beta2), float(eps),
                                     float(weight_decay), _stream()), "mmrec_adam_f32")


class _ProjectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, weight, bias, idx):
        ctx.save_for_backward(table, weight, idx)
        ctx.has_bias = bias is not None
        ctx.table_param = table if isinstance(table, torch.nn.Parameter) else None
        return project_raw(table, weight, bias, idx, False)

    @staticmethod
    def backward(ctx, g):
        # Backward of nn.Linear over the (gathered) table on the kernels of csrc/train.cu.
        table, weight, idx = ctx.saved_tensors
        g = _f32c(g)
        gw = gb = gt = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw, gb = linear_wgrad(g, table, idx, want_bias=ctx.has_bias and ctx.needs_input_grad[2])
            if not ctx.needs_input_grad[1]:
                gw = None
        if ctx.needs_input_grad[0]:
            G = g if idx is None else index_sum_rows(g, idx, table.shape[0])     # d-wide scatter; the table gradient is G @ W
            p = ctx.table_param
            owner = getattr(p, "_mmrec_defer", None) if p is not None else None
            if owner is not None and owner() is not None and getattr(p, "_mmrec_pending", None) is None \
                    and weight.shape[0] <= 128 and weight.shape[1] % 4 == 0:
                # The optimiser (optim.FusedAdam) asked for the gradient in factored form: it updates the table with G @ W
                # computed inside its kernel, so the [n_items, F] gradient never exists.  `.grad` stays None for this table.
                p._mmrec_pending = (G, weight, weight._version)
            else:
                gt = linear_dgrad(G, weight)
        return gt, gw, gb, None


def project(table, weight, bias=None, idx=None, l2_normalize=False) -> torch.Tensor:
    """`Linear(table)[idx]` computed only for the gathered rows (`src/models/freedom.py:205-209`,
    `bm3.py:102-104`, `mgcn.py:148-150`); `l2_normalize` adds `F.normalize` (`mmgcn.py:165-168`)."""
    needs_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (table, weight, bias))
    if not needs_grad:
        return project_raw(table, weight, bias, idx, l2_normalize)
    y = _ProjectFn.apply(table, weight, bias, idx)
    return torch.nn.functional.normalize(y) if l2_normalize else y


# -- a5b: MGCN's row-wise fusion (csrc/fuse.cu), inference form ---------------------------------------------------------
def gate_rows(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], mul: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`mul * sigmoid(x @ weight.T + bias)` in one kernel (`mmrec_gate_rows_f32`): MGCN's behaviour-guided purifier
    `item_id_embedding.weight * gate_v(image_feats)` (src/models/mgcn.py:153-154).  No autograd."""
    _need_cuda(x, weight, bias, mul, out)
    x, weight = _f32c(x), _f32c(weight)
    n, d = x.shape
    if weight.shape != (d, d):
        raise MMRecError(f"gate_rows: weight must be [{d}, {d}]")
    out = torch.empty(n, d, dtype=torch.float32, device=x.device) if out is None else out
    check(_lib.load().mmrec_gate_rows_f32(n, d, _ptr(x), _ptr(weight), _ptr(None if bias is None else _f32c(bias)),
                                          _ptr(None if mul is None else _f32c(mul)), _ptr(out), _stream()), "mmrec_gate_rows_f32")
    return out


def mgcn_fuse(img, txt, content, q_w, q_b, q_w2, gi_w, gi_b, gt_w, gt_b, want_side: bool = False):
    """MGCN's attention over the two modality views, preference gates and `content + side` (src/models/mgcn.py:187-201)
    for all rows in one kernel (`mmrec_mgcn_fuse_f32`).  Returns `all_embeds` (and `side` if asked).  No autograd."""
    _need_cuda(img, txt, content, q_w, q_b, q_w2, gi_w, gi_b, gt_w, gt_b)
    img, txt, content = _f32c(img), _f32c(txt), _f32c(content)
    n, d = img.shape
    if txt.shape != (n, d) or content.shape != (n, d):
        raise MMRecError("mgcn_fuse: img, txt and content must share one shape")
    out = torch.empty(n, d, dtype=torch.float32, device=img.device)
    side = torch.empty_like(out) if want_side else None
    check(_lib.load().mmrec_mgcn_fuse_f32(n, d, _ptr(img), _ptr(txt), _ptr(content), _ptr(_f32c(q_w)), _ptr(_f32c(q_b)),
                                          _ptr(_f32c(q_w2).reshape(-1)), _ptr(_f32c(gi_w)), _ptr(_f32c(gi_b)), _ptr(_f32c(gt_w)),
                                          _ptr(_f32c(gt_b)), _ptr(out), _ptr(side), _stream()), "mmrec_mgcn_fuse_f32")
    return (out, side) if want_side else out


# ------------------------------------------------------------------------------------------------
# K3: scoring, mask, top-k
# ------------------------------------------------------------------------------------------------
def set_score_path(path):
    """"simt" (0): exact fp32 CUDA cores; "tc" (1): tcgen05 3xTF32 + mask + radix top-k kernels; "auto" (2, default):
    fused where its shape rules allow, else tc; "fused" (3): tcgen05 with the top-k fused into the GEMM epilogue."""
    if isinstance(path, str):
        path = {"simt": 0, "tc": 1, "auto": 2, "fused": 3}[path]
    _lib.load().mmrec_score_set_path(int(path))


def score(user_e, item_e, users=None) -> torch.Tensor:
    """S = U[users] I^T, freshly allocated fp32 [B, n_items] owned by the caller (the trainer mutates it):
    `torch.matmul(u_embeddings, restore_item_e.transpose(0, 1))` (`src/models/freedom.py:216-220`)."""
    _need_cuda(user_e, item_e, users)
    lib = _lib.load()
    user_e, item_e = _f32c(user_e), _f32c(item_e)
    if users is not None:
        users = users.to(torch.int64).contiguous()
    B = user_e.shape[0] if users is None else users.numel()
    n_items, d = item_e.shape
    out = torch.empty(B, n_items, dtype=torch.float32, device=item_e.device)
    ws = _ws("score", lib.mmrec_score_workspace_bytes(B, n_items, d), item_e.device)
    check(lib.mmrec_score_f32(B, _ptr(users), _ptr(user_e), user_e.stride(0), n_items, _ptr(item_e), item_e.stride(0), d,
                              _ptr(out), n_items, _ptr(ws), ws.numel(), _stream()), "mmrec_score_f32")
    return out


def mask_topk(scores: torch.Tensor, mask: Optional[torch.Tensor], k: int, item_offset: int = 0):
    """`scores[mask[0], mask[1]] = -1e10; torch.topk(scores, k)` (`src/common/trainer.py:307-309`), in place on
    `scores`.  Returns (values, indices); equal scores come out in ascending item index."""
    _need_cuda(scores, mask)
    lib = _lib.load()
    if not (scores.is_contiguous() and scores.dtype == torch.float32 and scores.dim() == 2):
        raise MMRecError("mask_topk: scores must be contiguous float32 [B, n_items]")
    B, n_items = scores.shape
    if mask is not None and mask.numel() > 0:
        mask = mask.to(torch.int64).contiguous()
        check(lib.mmrec_mask_f32(mask.shape[1], _ptr(mask[0]), _ptr(mask[1]), B, n_items, item_offset, _ptr(scores),
                                 n_items, _stream()), "mmrec_mask_f32")
    idx = torch.empty(B, k, dtype=torch.int64, device=scores.device)
    val = torch.empty(B, k, dtype=torch.float32, device=scores.device)
    check(lib.mmrec_topk_rows_f32(B, n_items, _ptr(scores), n_items, k, item_offset, _ptr(idx), _ptr(val), _stream()),
          "mmrec_topk_rows_f32")
    return val, idx


class Catalog:
    """The item side of the scoring contraction prepared once per embedding table (`mmrec_catalog_pack_f32`): fp16
    operand tiles (power-of-two scaled, 11 significand bits) + the maximum row norm of the error bound.  The reference re-reads the same `restore_item_e` for every
    evaluation batch (`src/common/trainer.py:302-310`); a model keeps one Catalog next to its cached evaluation
    embeddings and drops it with them.  Holds a reference to `item_e`: the pair must stay consistent."""

    def __init__(self, item_e: torch.Tensor):
        _need_cuda(item_e)
        lib = _lib.load()
        self.item_e = _f32c(item_e)
        n_items, d = self.item_e.shape
        nbytes = lib.mmrec_catalog_bytes(n_items, d)
        if nbytes == 0:
            raise MMRecError(f"catalog: no tensor-core path for d = {d}")
        self.buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=item_e.device)
        self.ptr = (self.buf.data_ptr() + 1023) // 1024 * 1024
        check(lib.mmrec_catalog_pack_f32(n_items, _ptr(self.item_e), self.item_e.stride(0), d, self.ptr, nbytes, _stream()),
              "mmrec_catalog_pack_f32")


def score_topk(user_e, item_e, users, mask, k: int, item_offset: int = 0, out=None, catalog: Optional[Catalog] = None):
    """Fused `full_sort_predict` + mask + top-k (`src/models/freedom.py:216-220` + `src/common/trainer.py:304-309`)
    without materialising the [B, n_items] score matrix in HBM.  Returns (values [B,k], indices int64 [B,k]); `out` =
    (values, indices) buffers to write into (e.g. peer-mapped memory in the sharded evaluation); `catalog` = the
    `Catalog` of exactly this `item_e` (else the item operand is packed inside the call)."""
    _need_cuda(user_e, item_e, users, mask)
    lib = _lib.load()
    user_e, item_e = _f32c(user_e), _f32c(item_e)
    if catalog is not None and (catalog.item_e.data_ptr() != item_e.data_ptr() or catalog.item_e.shape != item_e.shape):
        raise MMRecError("score_topk: the catalog was packed from a different item table")
    if users is not None:
        users = users.to(torch.int64).contiguous()
    B = user_e.shape[0] if users is None else users.numel()
    n_items, d = item_e.shape
    if out is not None:
        val, idx = out
        assert val.shape == (B, k) and idx.shape == (B, k) and val.dtype == torch.float32 and idx.dtype == torch.int64 \
            and val.is_contiguous() and idx.is_contiguous()
    else:
        idx = torch.empty(B, k, dtype=torch.int64, device=item_e.device)
        val = torch.empty(B, k, dtype=torch.float32, device=item_e.device)
    m0 = m1 = None
    nnz = 0
    if mask is not None and mask.numel() > 0:
        mask = mask.to(torch.int64).contiguous()
        m0, m1, nnz = mask[0], mask[1], mask.shape[1]
    nbytes = lib.mmrec_score_topk_workspace_bytes(B, n_items, d, k) + 4 * nnz + 4096
    ws = _ws("score_topk", nbytes, item_e.device)
    _last_fused.update(ws=ws, args=(B, n_items, d, k, nnz, int(catalog is None)))
    check(lib.mmrec_score_topk_cat_f32(B, _ptr(users), _ptr(user_e), user_e.stride(0), n_items, _ptr(item_e),
                                       item_e.stride(0), d, None if catalog is None else catalog.ptr, nnz, _ptr(m0), _ptr(m1), k,
                                       item_offset, _ptr(idx), _ptr(val), _ptr(ws), ws.numel(), _stream()), "mmrec_score_topk_cat_f32")
    return val, idx


_last_fused: dict = {}


def fused_fallback_rows(*_ignored) -> int:
    """Diagnostic (synchronises): how many rows of the last row block of the last fused score_topk call went through the
    exact fp32 kernel; -1 when that call did not take the fused path."""
    if not _last_fused:
        return -1
    return int(_lib.load().mmrec_debug_fused_fallback_rows(_ptr(_last_fused["ws"]), *_last_fused["args"]))


def fused_stage_times():
    """Tuning aid: device microseconds of the stages of the last fused score_topk call (env MMREC_CF_TIMING must be set
    before the first call): [catalogue pack, prep + mask, pass 1, threshold, pass 2, finalists, exact rows]."""
    import ctypes
    buf = (ctypes.c_float * 16)()
    n = _lib.load().mmrec_debug_cf_timing(ctypes.cast(buf, ctypes.c_void_p), 16)
    return [float(buf[i]) for i in range(n)]


def topk_merge(vals: torch.Tensor, idx: torch.Tensor):
    """Merge per-shard top-k lists [parts, B, k] into the global top-k [B, k] (SURVEY.md 8e eval collective)."""
    _need_cuda(vals, idx)
    lib = _lib.load()
    vals, idx = _f32c(vals), idx.to(torch.int64).contiguous()
    parts, B, k = vals.shape
    out_i = torch.empty(B, k, dtype=torch.int64, device=vals.device)
    out_v = torch.empty(B, k, dtype=torch.float32, device=vals.device)
    check(lib.mmrec_topk_merge(parts, B, k, _ptr(vals), _ptr(idx), _ptr(out_i), _ptr(out_v), _stream()), "mmrec_topk_merge")
    return out_v, out_i


def _ptr_array(ptrs):
    import ctypes
    arr = (ctypes.c_void_p * len(ptrs))(*[int(x) for x in ptrs])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def topk_merge_peers(val_ptrs, idx_ptrs, B, k, device, idx_mul=1, idx_add=0, row0=0, n_rows=None, sync=None):
    """`topk_merge` over lists left where each rank wrote them: raw device addresses of `parts` [B, k] value (fp32) and
    index (int64) lists, rank order; index -> idx * idx_mul + part * idx_add (round-robin shards: world, 1).  Rows
    [row0, row0 + n_rows) only (default: all): returns ([n_rows, k] values, indices).  `sync` = (flag_ptrs, state, rank):
    the ranks are synchronised inside the kernel (before the lists are read) instead of by a barrier launch."""
    lib = _lib.load()
    n_rows = B - row0 if n_rows is None else n_rows
    va, vap = _ptr_array(val_ptrs)
    ia, iap = _ptr_array(idx_ptrs)
    out_i = torch.empty(n_rows, k, dtype=torch.int64, device=device)
    out_v = torch.empty(n_rows, k, dtype=torch.float32, device=device)
    fa = fap = None
    if sync is not None:
        fa, fap = _ptr_array(sync[0])
    check(lib.mmrec_topk_merge_peers(len(val_ptrs), B, k, vap, iap, int(idx_mul), int(idx_add), int(row0), int(n_rows), _ptr(out_i),
                                     _ptr(out_v), fap, None if sync is None else _ptr(sync[1]), 0 if sync is None else int(sync[2]),
                                     _stream()), "mmrec_topk_merge_peers")
    return out_v, out_i


def peer_sum(part_ptrs, n, acc_in=None, acc_out=None, acc_div=1.0, sum_out=None):
    """K4: `sum_out = sum_r parts[r]` (rank order), `acc_out = (acc_in + sum) / acc_div` -- the user-embedding exchange of
    the item-sharded propagation as one pass over peer-mapped buffers (`part_ptrs`: raw device addresses, rank order)."""
    lib = _lib.load()
    arr, arrp = _ptr_array(part_ptrs)
    check(lib.mmrec_peer_sum_f32(int(n), len(part_ptrs), arrp, _ptr(acc_in), _ptr(acc_out), float(acc_div),
                                 _ptr(sum_out), _stream()), "mmrec_peer_sum_f32")
    return sum_out, acc_out


def peer_reduce_push(part_ptrs, dst_ptrs, n, rank, acc_in=None, acc_out=None, acc_div=1.0, final_layer=False):
    """K4, reduce-scatter + all-gather form (`mmrec_peer_reduce_push_f32`): this rank sums ITS slice of all partials and
    stores the result (final layer: `(acc + sum) / acc_div`) into that slice of every rank's destination buffer;
    `acc_in` / `acc_out` hold this rank's slice of the running layer sum."""
    lib = _lib.load()
    pa, pap = _ptr_array(part_ptrs)
    da, dap = _ptr_array(dst_ptrs)
    check(lib.mmrec_peer_reduce_push_f32(int(n), len(part_ptrs), int(rank), pap, dap, _ptr(acc_in), _ptr(acc_out), float(acc_div),
                                         int(bool(final_layer)), _stream()), "mmrec_peer_reduce_push_f32")


def peer_exchange(part_ptrs, dst_ptrs, flag_ptrs, state, n, rank, acc_in=None, acc_out=None, acc_div=1.0, final_layer=False):
    """`peer_reduce_push` with both rank synchronisations inside the kernel (`mmrec_peer_exchange_f32`): one launch per layer."""
    lib = _lib.load()
    pa, pap = _ptr_array(part_ptrs)
    da, dap = _ptr_array(dst_ptrs)
    fa, fap = _ptr_array(flag_ptrs)
    check(lib.mmrec_peer_exchange_f32(int(n), len(part_ptrs), int(rank), pap, dap, fap, _ptr(state), _ptr(acc_in), _ptr(acc_out),
                                      float(acc_div), int(bool(final_layer)), _stream()), "mmrec_peer_exchange_f32")


def peer_barrier(flag_ptrs, state, rank):
    """Device-side barrier of the ranks on the current stream over the same flags (`mmrec_peer_barrier`)."""
    fa, fap = _ptr_array(flag_ptrs)
    check(_lib.load().mmrec_peer_barrier(len(flag_ptrs), int(rank), fap, _ptr(state), _stream()), "mmrec_peer_barrier")


def peer_gather(src_ptrs, n_each, dst: torch.Tensor):
    """`dst[p * n_each + i] = src[p][i]`: all-gather of a sharded table by peer loads (`mmrec_peer_gather_f32`)."""
    lib = _lib.load()
    sa, sap = _ptr_array(src_ptrs)
    check(lib.mmrec_peer_gather_f32(int(n_each), len(src_ptrs), sap, _ptr(dst), _stream()), "mmrec_peer_gather_f32")
    return dst


def topk_metric_sums(topk_idx: torch.Tensor, pos_ptr: torch.Tensor, pos_items: torch.Tensor, disc: torch.Tensor, idcg_all: torch.Tensor,
                     sums: torch.Tensor):
    """f2: adds, per position j < K, the sum over the rows of `topk_idx` of recall / ndcg / precision / map at j + 1 to `sums`
    [4, K] float64 (`mmrec_topk_metrics_f64`; src/utils/topk_evaluator.py:70-102 + src/utils/metrics.py:12-105 on the device)."""
    _need_cuda(topk_idx, pos_ptr, pos_items, disc, idcg_all, sums)
    lib = _lib.load()
    n, K = topk_idx.shape
    assert topk_idx.dtype == torch.int64 and topk_idx.is_contiguous() and pos_ptr.dtype == torch.int64 and pos_items.dtype == torch.int64
    assert sums.dtype == torch.float64 and sums.shape == (4, K) and sums.is_contiguous() and disc.dtype == torch.float64
    check(lib.mmrec_topk_metrics_f64(n, K, _ptr(topk_idx), _ptr(pos_ptr), _ptr(pos_items), _ptr(disc), _ptr(idcg_all), _ptr(sums),
                                     _stream()), "mmrec_topk_metrics_f64")
    return sums


def bipartite_norm(users: torch.Tensor, items: torch.Tensor, n_users: int, n_items: int, eps: float = 1e-7) -> torch.Tensor:
    """fp32 1/sqrt((d_u+eps)(d_i+eps)) per edge, as `_normalize_adj_m` (`src/models/freedom.py:145-154`)."""
    _need_cuda(users, items)
    lib = _lib.load()
    users, items = users.to(torch.int64).contiguous(), items.to(torch.int64).contiguous()
    vals = torch.empty(users.numel(), dtype=torch.float32, device=users.device)
    ws = _ws("bnorm", lib.mmrec_bipartite_norm_workspace_bytes(n_users, n_items), users.device)
    check(lib.mmrec_bipartite_norm_f32(users.numel(), _ptr(users), _ptr(items), n_users, n_items, eps, _ptr(vals), _ptr(ws),
                                       ws.numel(), _stream()), "mmrec_bipartite_norm_f32")
    return vals
