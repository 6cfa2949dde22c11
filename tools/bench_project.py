"""Micro-benchmark of the modality projection (gather -> linear -> optional L2 norm), device time per call.
    python tools/bench_project.py [--rows 7050] [--F 4096,4128,384] [--d 64] [--reps 10]
Non-power-of-two F values probe how much of the time is HBM bank conflicts on the 16 KB row pitch."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmrec_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=7050); ap.add_argument("--F", default="4096,4128,4352,384")
ap.add_argument("--d", type=int, default=64); ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.isfile("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for F in [int(x) for x in a.F.split(",")]:
    g = torch.Generator(device=dev); g.manual_seed(F)
    table = torch.randn(a.rows, F, device=dev, generator=g)
    W = torch.randn(a.d, F, device=dev, generator=g) * 0.02
    b = torch.randn(a.d, device=dev, generator=g)
    ts = []
    for r in range(a.reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = ops.project(table, W, b); e1.record(); torch.cuda.synchronize()
        if r >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    ref = torch.nn.functional.linear(table.double(), W.double(), b.double())
    err = ((y.double() - ref).abs().max() / ref.abs().max()).item()
    byt = 4.0 * (a.rows * F + a.d * F + a.rows * a.d)
    print(f"F={F:5d} rows={a.rows} d={a.d}: {us:8.1f} us  {byt/us/1e3:7.1f} GB/s  frac={byt/us/1e3/peak:.3f}  max rel err {err:.2e}")
