"""Micro-benchmark of full-catalog scoring + mask + top-k (device time per 4096-user batch) for the three paths.
    python tools/bench_score.py [--workload baby] [--batch 4096] [--reps 10]"""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmrec_b200 import ops
from mmrec_b200.utils import synth
import bench as B

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="baby"); ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--reps", type=int, default=10); ap.add_argument("--paths", default="auto,fused,tc,simt")
a = ap.parse_args()
dev = torch.device("cuda:0")
wl = B.Workload(a.workload)
ue = torch.from_numpy(wl.user_emb).to(dev); ie = torch.from_numpy(wl.item_emb).to(dev)
nb = min(a.batch, wl.U)
users = torch.arange(nb, device=dev)
mask = torch.from_numpy(wl.eval_mask(0, nb)).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
peak = (json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"] if os.path.isfile("MEASURED_PEAKS.json") else 1590.0) / 2
flops = 2.0 * nb * wl.I * wl.d
print(f"{a.workload}: B={nb} I={wl.I} d={wl.d} mask_nnz={mask.shape[1]}  useful {flops/1e9:.2f} GFLOP per batch; tf32 peak {peak:.0f} TF/s")
ref = None
side = torch.cuda.Stream()
for path in a.paths.split(","):
    ops.set_score_path(path)
    ts = []
    for r in range(a.reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); val, idx = ops.score_topk(ue, ie, users, mask, 50); e1.record(); torch.cuda.synchronize()
        if r >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    if path in ("auto", "fused") and not os.environ.get("MMREC_CF_TIMING"):
        # the same call replayed from a CUDA graph (no host launch cost), item operand packed once (ops.Catalog)
        cat = ops.Catalog(ie)
        with torch.cuda.stream(side):
            ops.score_topk(ue, ie, users, mask, 50, catalog=cat); torch.cuda.synchronize()
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=side):
                gv, gi = ops.score_topk(ue, ie, users, mask, 50, catalog=cat)
        torch.cuda.synchronize()
        gts = []
        for r in range(a.reps + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gph.replay(); e1.record(); torch.cuda.synchronize()
            if r >= 3: gts.append(e0.elapsed_time(e1) * 1e3)
        print(f"   graph replay with a packed catalogue: {float(np.median(gts)):.1f} us/batch (identical result: {bool(torch.equal(gi, idx))})")
    fb = ops.fused_fallback_rows() if path in ("fused", "auto") else 0
    agree = "" if ref is None else f" rows identical to {refname}: {(idx == ref).all(dim=1).float().mean().item():.4f}"
    if ref is None: ref, refname = idx, path
    if os.environ.get("MMREC_CF_TIMING") and path in ("auto", "fused"):
        print("   stages us [pack, prep+mask, pass1, thr, pass2, final, exact]:", [round(x, 1) for x in ops.fused_stage_times()])
    print(f"{path:6s} {us:9.1f} us/batch  {nb*wl.I/us/1e3:8.2f} G items/s  {flops/us/1e6:7.2f} TFLOP/s useful  frac={flops/us/1e6/peak:.3f}  fallback_rows={fb}{agree}")
ops.set_score_path("auto")
