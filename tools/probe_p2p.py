"""NVLink peer-memory bandwidth of this library's exchange kernels (run under torchrun, 2+ ranks):
    python -m torch.distributed.run --nproc-per-node 2 tools/probe_p2p.py [--mb 128]
gather = reads only (every rank reads the other ranks' block), reduce_push = (N-1)/N read + (N-1)/N written."""
import argparse, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmrec_b200 import ops

ap = argparse.ArgumentParser(); ap.add_argument("--mb", type=int, default=128); a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
n = a.mb * (1 << 20) // 4
buf = symm.empty(2 * n, dtype=torch.float32, device=dev); hdl = symm.rendezvous(buf, dist.group.WORLD)
buf.normal_(); hdl.barrier(channel=0); torch.cuda.synchronize()
src_ptrs = [int(p) for p in hdl.buffer_ptrs]; dst_ptrs = [int(p) + 4 * n for p in hdl.buffer_ptrs]
out = torch.empty(world * (n // world), dtype=torch.float32, device=dev)
acc = torch.zeros((n // 4 + world - 1) // world * 4, dtype=torch.float32, device=dev)

def timed(fn, reps=5):
    ts = []
    for r in range(reps + 2):
        hdl.barrier(channel=1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if r >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))

each = n // world
t_g = timed(lambda: ops.peer_gather(src_ptrs, each, out))
t_r = timed(lambda: ops.peer_reduce_push(src_ptrs, dst_ptrs, n, rank, acc_in=None, acc_out=None))
if rank == 0:
    rem_g = (world - 1) * each * 4
    rem_r = (world - 1) / world * n * 4
    print(f"world {world}, {a.mb} MB table: gather {t_g:.1f} us = {rem_g / t_g / 1e3:.0f} GB/s of peer reads per rank; "
          f"reduce_push {t_r:.1f} us = {rem_r / t_r / 1e3:.0f} GB/s read + {rem_r / t_r / 1e3:.0f} GB/s written per rank")
dist.barrier(); dist.destroy_process_group()
