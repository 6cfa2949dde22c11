"""Micro-benchmark of the f1 kernels (csrc/train.cu) beside the torch / cuBLAS sequence they replace; device time per call,
L2 flushed before every call.
    python tools/bench_train.py [--rows 7000] [--F 4096,384] [--d 64] [--batch 4096] [--reps 8]
Prints one JSON line per F (also used for profiles/r02_train.md)."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmrec_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=7000); ap.add_argument("--F", default="4096,384")
ap.add_argument("--d", type=int, default=64); ap.add_argument("--batch", type=int, default=4096); ap.add_argument("--reps", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda:0")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.isfile("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    ts = []
    for r in range(a.reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if r >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


for F in [int(x) for x in a.F.split(",")]:
    g = torch.Generator(device=dev); g.manual_seed(F)
    table = torch.randn(a.rows, F, device=dev, generator=g)
    W = torch.randn(a.d, F, device=dev, generator=g) * 0.02
    up = torch.randn(a.rows, a.d, device=dev, generator=g) * 0.01
    idx = torch.randint(0, a.rows, (a.batch,), device=dev, generator=g)
    upb = torch.randn(a.batch, a.d, device=dev, generator=g) * 0.01
    m, v = torch.zeros_like(table), torch.zeros_like(table)
    out = {"F": F, "rows": a.rows, "d": a.d, "batch": a.batch}
    tb = 4.0 * a.rows * F
    # ours
    out["wgrad_us"] = timed(lambda: ops.linear_wgrad(up, table))
    out["wgrad_gather_us"] = timed(lambda: ops.linear_wgrad(upb, table, idx))
    out["index_sum_rows_us"] = timed(lambda: ops.index_sum_rows(upb, idx, a.rows))
    out["dgrad_store_us"] = timed(lambda: ops.linear_dgrad(up, W))
    out["dgrad_adam_us"] = timed(lambda: ops.linear_dgrad_adam(up, W, table, m, v, 0.9, 0.999, 1e-8, 0.0, -1e-3, 0.0316))
    out["dgrad_adam_gbs"] = 6 * tb / out["dgrad_adam_us"] / 1e3
    out["dgrad_adam_hbm_frac"] = out["dgrad_adam_gbs"] / peak
    grad = torch.randn_like(table)
    out["adam_plain_us"] = timed(lambda: ops.adam_step([(table, grad, m, v, -1e-3, 0.0316)], 0.9, 0.999, 1e-8, 0.0))
    out["adam_plain_gbs"] = 7 * tb / out["adam_plain_us"] / 1e3
    # torch / cuBLAS: what the reference's step does for this table (src/models/freedom.py:205-209 backward + trainer.py:189)
    out["torch_wgrad_us"] = timed(lambda: up.t().mm(table))
    out["torch_dgrad_us"] = timed(lambda: up.mm(W))
    p = torch.nn.Parameter(table.clone()); p.grad = grad
    opt = torch.optim.Adam([p], lr=1e-3)
    opt.step()
    out["torch_adam_us"] = timed(lambda: opt.step())
    out["torch_index_add_us"] = timed(lambda: torch.zeros_like(table).index_add_(0, idx, upb.mm(W)))
    out["ours_dense_table_step_us"] = out["wgrad_us"] + out["dgrad_store_us"] + out["adam_plain_us"]      # FusedAdam default
    out["ours_factored_table_step_us"] = out["wgrad_us"] + out["dgrad_adam_us"]                           # factored=True
    out["dgrad_cluster"] = os.environ.get("MMREC_DGRAD_CLUSTER", "0")
    out["torch_table_step_us"] = out["torch_wgrad_us"] + out["torch_dgrad_us"] + out["torch_adam_us"]
    print(json.dumps({k: (round(x, 2) if isinstance(x, float) else x) for k, x in out.items()}))
