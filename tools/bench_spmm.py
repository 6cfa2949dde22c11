"""Micro-benchmark of the SpMM kernel alone (device time, CUDA events, L2 flushed between repetitions).
    python tools/bench_spmm.py [--workload baby] [--layers 3] [--reps 20]
Prints one line per lanes-per-row setting: us per layer, achieved algorithmic GB/s, fraction of measured HBM peak."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmrec_b200 import _lib, graph, ops
from mmrec_b200.utils import synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="baby"); ap.add_argument("--layers", type=int, default=3)
ap.add_argument("--reps", type=int, default=20); ap.add_argument("--d", type=int, default=0)
ap.add_argument("--flush", default="write", choices=["none", "write", "read", "write+read"])
ap.add_argument("--uniform", action="store_true", help="uniform item popularity (no heavy rows): isolates the cost of the power-law tail")
ap.add_argument("--users", type=int, default=0); ap.add_argument("--items", type=int, default=0); ap.add_argument("--edges", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
U, I, E, d, _ = synth.SHAPES[a.workload]
U, I, E, d = a.users or U, a.items or I, a.edges or E, a.d or d
if a.uniform:
    rng = np.random.default_rng(0)
    key = np.unique(rng.integers(0, U, int(E * 1.02)) * I + rng.integers(0, I, int(E * 1.02)))
    tu, ti = key // I, key % I
else:
    g = synth.make_graph(U, I, E, 0)
    tu, ti = g.train
n = U + I
rows_, cols_, vals_ = graph.norm_adj_entries(tu, ti, U, I)
def build(seg, light=None):
    return ops.CSR.from_coo(torch.from_numpy(rows_).to(dev), torch.from_numpy(cols_).to(dev), torch.from_numpy(vals_).to(dev),
                            n, n, sum_duplicates=False, symmetric=True, seg=seg, light_max=light)
adj = build(ops.SEG)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.isfile("MEASURED_PEAKS.json") else 6650.0
ego = torch.randn(n, d, device=dev) * 0.1
_fbuf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
_rbuf = torch.ones(64 << 20, dtype=torch.float32, device=dev)
class _Flush:
    def zero_(self):
        if "write" in a.flush: _fbuf.zero_()
        if "read" in a.flush: _rbuf.sum()
flush = _Flush()
lib = _lib.load()
print(f"graph {a.workload}: N={n} nnz={adj.nnz} d={d} tasks={adj.n_tasks} split_rows={adj.n_split} longest={adj.longest_row} "
      f"bytes/layer={adj.algorithmic_bytes(d)/1e6:.2f} MB")
for seg, light, lanes in [(s_, lt_, l_) for s_ in (512,) for lt_ in (16, 32) for l_ in (8, 16)]:
    if lanes and (d % (4 * lanes) or d // (4 * lanes) > 4):
        continue
    adj = build(seg, light)
    lib.mmrec_spmm_set_lanes(lanes)
    for plan in (True,):
        out = None
        ts = []
        for r in range(a.reps + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if plan:
                out = ops.propagate_mean(adj, ego, a.layers)
            else:
                acc = torch.empty_like(ego); x = ego
                for l in range(1, a.layers + 1):
                    y = torch.empty_like(ego)
                    ops.spmm_raw(adj, x, Y=y, acc_in=ego if l == 1 else acc, acc_out=acc, use_plan=False); x = y
            e1.record(); torch.cuda.synchronize()
            if r >= 3: ts.append(e0.elapsed_time(e1) * 1e3 / a.layers)
        us = float(np.median(ts))
        gbs = adj.algorithmic_bytes(d) / us / 1e3
        print(f"seg={seg:4d} light<={light:2d} tasks={adj.n_tasks:6d} cta_tasks={adj.n_cta_tasks:5d} lanes={lanes:2d} plan={int(plan)}  {us:8.2f} us/layer  {gbs:8.1f} GB/s  frac={gbs/peak:.3f}  edges/s={adj.nnz/us*1e6:.3e}")
lib.mmrec_spmm_set_lanes(0)
adj = build(ops.SEG)
# graph replay (no host launch gaps)
gph = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s), torch.no_grad():
    ops.propagate_mean(adj, ego, a.layers)
    torch.cuda.synchronize()
    with torch.cuda.graph(gph, stream=s):
        out = ops.propagate_mean(adj, ego, a.layers)
ts = []
for r in range(a.reps + 3):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gph.replay(); e1.record(); torch.cuda.synchronize()
    if r >= 3: ts.append(e0.elapsed_time(e1) * 1e3 / a.layers)
us = float(np.median(ts)); gbs = adj.algorithmic_bytes(d) / us / 1e3
print(f"cuda-graph replay (default lanes): {us:8.2f} us/layer  {gbs:8.1f} GB/s  frac={gbs/peak:.3f}")
