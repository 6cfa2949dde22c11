"""Parity at BASELINE.json's configuration sizes (configs[1..3]: baby / sports / clothing graphs, d = 64) and at the
per-GPU shape of configs[4] (d = 128, >= 125k items): the CUDA path against the CPU oracle on the same seeded inputs.

Propagation: the oracle's `torch.sparse.mm` chain (src/models/freedom.py:164-178) on the full graph, 1e-5 relative
(bar 1e-4).  Projection: `Linear` over the first rows of a 4096-d feature table.  Fused scoring + mask + top-50: one
evaluation batch of 4096 users against an fp64 re-score of the same fp32 embeddings (near-tie rule).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mmrec_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmrec_b200 import _lib
    _lib.require_device()
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def near_tie_check(idx, ref, ri, scale, tol=4e-6):
    got = idx.cpu().numpy()
    bad = np.nonzero((got != ri).any(axis=1))[0]
    for b in bad:
        cols = np.nonzero(got[b] != ri[b])[0]
        gap = np.abs(ref[b, got[b, cols]].numpy() - ref[b, ri[b, cols]].numpy()).max()
        assert gap < tol * scale, f"row {b}: non-tie mismatch, gap {gap}"
    return len(bad)


@pytest.mark.parametrize("name,n_layers", [("baby", 3), ("sports", 2), ("clothing", 2)])
def test_config_size_propagation_and_topk_vs_oracle(dev, name, n_layers):
    from mmrec_b200 import graph, ops
    from mmrec_b200.utils import synth
    g = synth.named(name)
    U, I = g.n_users, g.n_items
    tu, ti = g.train
    gen = torch.Generator().manual_seed(1)
    bound_u, bound_i = (6.0 / (U + 64)) ** 0.5, (6.0 / (I + 64)) ** 0.5               # xavier_uniform, freedom.py:51-52
    ego = torch.cat([(torch.rand(U, 64, generator=gen) * 2 - 1) * bound_u, (torch.rand(I, 64, generator=gen) * 2 - 1) * bound_i])
    # ---- propagation, full graph
    A = graph.build_norm_adj((tu, ti), U, I, dev)
    out = ops.propagate_mean(A, ego.to(dev), n_layers)
    ref = O.propagate_mean(O.norm_adj_coo(tu, ti, U, I), ego, n_layers)
    assert rel(out, ref) < 1e-5
    err = (out.cpu() - ref).abs().max().item()
    assert err < 1e-6 + 1e-4 * ref.abs().max().item()
    # ---- fused scoring + mask + top-50, one evaluation batch, on the propagated embeddings
    u_g, i_g = out[:U].contiguous(), out[U:].contiguous()
    lo, hi = 4096, 8192
    users = torch.arange(lo, hi)
    sel = (tu >= lo) & (tu < hi)
    order = np.argsort(tu[sel], kind="stable")
    mask = torch.from_numpy(np.stack([tu[sel][order] - lo, ti[sel][order]]))
    ops.set_score_path("fused")
    try:
        val, idx = ops.score_topk(u_g, i_g, users.to(dev), mask.to(dev), 50)
        assert ops.fused_fallback_rows() <= 4                                       # (heavy users only)
        cat = ops.Catalog(i_g)
        val_c, idx_c = ops.score_topk(u_g, i_g, users.to(dev), mask.to(dev), 50, catalog=cat)
        assert torch.equal(idx, idx_c) and torch.equal(val, val_c)
    finally:
        ops.set_score_path("auto")
    sref = O.full_sort_scores(u_g.cpu().double(), i_g.cpu().double(), users)
    sref[mask[0], mask[1]] = -1e10
    rv, ri = O.topk_tie_low_index(sref.numpy(), 50)
    scale = sref[sref > -1e9].abs().max().item()
    n_bad = near_tie_check(idx, sref, ri, scale)
    assert n_bad <= 4096 // 50                                                      # fp32 vs fp64 order: a handful of near ties
    assert (val.cpu().double() - torch.from_numpy(rv)).abs().max().item() < 4e-6 * scale
    # Recall@20-style agreement: the top-20 SETS agree on every row that has no near tie at rank 20
    same20 = (np.sort(idx.cpu().numpy()[:, :20], axis=1) == np.sort(ri[:, :20], axis=1)).all(axis=1).mean()
    assert same20 > 0.995


@pytest.mark.parametrize("name", ["sports", "clothing"])
def test_config_size_projection_vs_oracle(dev, name):
    from mmrec_b200 import ops
    from mmrec_b200.utils import synth
    _, I, _, d, F = synth.SHAPES[name]
    gen = torch.Generator().manual_seed(2)
    X = torch.randn(I, F, generator=gen)
    W = torch.randn(d, F, generator=gen) / F ** 0.5
    b = torch.randn(d, generator=gen) * 0.1
    y = ops.project(X.to(dev), W.to(dev), b.to(dev))
    assert rel(y, O.project(X, W, b)) < 5e-5                                        # bar: 1e-4 (3xTF32 over F = 4096 terms)
    idx = torch.randint(0, I, (2048,), generator=gen)                               # the calculate_loss gather (freedom.py:205-209)
    assert rel(ops.project(X.to(dev), W.to(dev), b.to(dev), idx=idx.to(dev)), O.project(X, W, b, idx=idx)) < 5e-5


def test_fused_topk_config5_shard_shape(dev):
    """One GPU's share of configs[4]: d = 128, 125k+ items (groups of 128 items), users in batches; no row may need the
    exact kernel, results against an fp64 re-score."""
    from mmrec_b200 import ops
    gen = torch.Generator().manual_seed(5)
    B, U, I, d, k = 384, 3000, 125_000 + 37, 128, 50
    ue = torch.randn(U, d, generator=gen) * 0.05
    ie = torch.randn(I, d, generator=gen) * 0.05
    ie[:2000] *= 3.0                                                                # a popular head with larger norms
    users = torch.randint(0, U, (B,), generator=gen)
    nm = B * 25
    mask = torch.stack([torch.randint(0, B, (nm,), generator=gen).sort().values, torch.randint(0, I, (nm,), generator=gen)])
    ops.set_score_path("fused")
    try:
        gmask = torch.stack([mask[0], mask[1] + 1000])                              # mask columns are GLOBAL item ids
        val, idx = ops.score_topk(ue.to(dev), ie.to(dev), users.to(dev), gmask.to(dev), k, item_offset=1000)
        assert ops.fused_fallback_rows() == 0
    finally:
        ops.set_score_path("auto")
    sref = O.full_sort_scores(ue.double(), ie.double(), users)
    sref[mask[0], mask[1]] = -1e10
    rv, ri = O.topk_tie_low_index(sref.numpy(), k)
    scale = sref[sref > -1e9].abs().max().item()
    near_tie_check(idx - 1000, sref, ri, scale)
    assert (val.cpu().double() - torch.from_numpy(rv)).abs().max().item() < 4e-6 * scale
