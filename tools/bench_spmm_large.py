"""SpMM on a graph whose dense operand does not fit the L2: plain CSR against the column-panelled form (ops.PanelCSR).
    python tools/bench_spmm_large.py [--users 1000000 --items 300000 --edges 16000000 --d 64] [--panel-mb 48]
Prints us per layer, algorithmic GB/s (SURVEY.md 8d bytes) and the fraction of the measured HBM peak."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmrec_b200 import graph, ops

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=1000000); ap.add_argument("--items", type=int, default=300000)
ap.add_argument("--edges", type=int, default=16000000); ap.add_argument("--d", type=int, default=64)
ap.add_argument("--panel-mb", type=int, default=48); ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
U, I, E, d = a.users, a.items, a.edges, a.d
w = (np.arange(I, dtype=np.float64) + 1.0) ** -0.8
cdf = np.cumsum(w); cdf /= cdf[-1]
tu = rng.integers(0, U, E); ti = np.minimum(np.searchsorted(cdf, rng.random(E)), I - 1)
key = np.unique(tu * I + ti); tu, ti = key // I, key % I
n = U + I
rows, cols, vals = graph.norm_adj_entries(tu, ti, U, I)
r, c, v = torch.from_numpy(rows).to(dev), torch.from_numpy(cols).to(dev), torch.from_numpy(vals).to(dev)
A = ops.CSR.from_coo(r, c, v, n, n, sum_duplicates=False, symmetric=True)
P = ops.PanelCSR.from_coo(r, c, v, n, n, d, panel_bytes=a.panel_mb << 20, sum_duplicates=False, symmetric=True)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.isfile("MEASURED_PEAKS.json") else 6650.0
X = torch.randn(n, d, device=dev) * 0.1
Y1, Y2 = torch.empty_like(X), torch.empty_like(X)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
print(f"N={n} nnz={A.nnz} d={d}: X {n*d*4/1e6:.0f} MB, algorithmic bytes/layer {A.algorithmic_bytes(d)/1e6:.0f} MB, {len(P.panels)} panels of <= {a.panel_mb} MB")
for name, M, Y in (("plain CSR", A, Y1), ("PanelCSR", P, Y2)):
    ts = []
    for rep in range(a.reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.spmm_raw(M, X, Y=Y); e1.record(); torch.cuda.synchronize()
        if rep >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    gbs = A.algorithmic_bytes(d) / us / 1e3
    print(f"{name:10s} {us:9.1f} us/layer  {A.nnz/us/1e3:6.2f} G nnz/s  {gbs:7.1f} GB/s algorithmic  frac={gbs/peak:.3f}")
print("max rel diff panel vs plain:", ((Y1 - Y2).norm() / Y1.norm()).item())
