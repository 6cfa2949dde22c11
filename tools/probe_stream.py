"""What does HBM deliver for K2's access pattern as a function of the burst length per row?
    python tools/probe_stream.py [--rows 7000] [--F 4096]"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmrec_b200 import _lib
ap = argparse.ArgumentParser(); ap.add_argument("--rows", type=int, default=7000); ap.add_argument("--F", type=int, default=4096)
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = _lib.load()
t = torch.randn(a.rows, a.F, device=dev); sink = torch.zeros(1, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for per_sm in (1, 2, 4):
    for R, burst in ((128, 512), (64, 1024), (32, 2048), (16, 4096), (4, 16384)):
        ts = []
        for r in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rc = lib.mmrec_debug_stream_probe(t.data_ptr(), a.rows, a.F, R, burst, per_sm, sink.data_ptr(), st); e1.record()
            torch.cuda.synchronize(); assert rc == 0
            if r >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
        us = float(np.median(ts))
        print(f"CTAs/SM {per_sm}  R={R:4d} rows x {burst:6d} B bursts: {us:7.1f} us  {a.rows * a.F * 4 / us / 1e3:7.1f} GB/s")
