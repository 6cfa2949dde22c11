"""Worker of tests/test_gpu_sharded.py (one process per GPU, launched by torch.distributed.run): the peer-memory sharded
path on real GPUs against the single-GPU kernels.  Prints one JSON line on rank 0 and exits non-zero on a mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from mmrec_b200 import ops, sharded
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    wl = bench.Workload(os.environ.get("MMREC_TEST_WORKLOAD", "small"), n_layers=3, items_scale=world)
    U, I, d, k = wl.U, wl.I, wl.d, 50
    shard = sharded.ItemShard(wl.tr_u, wl.tr_i, U, I, rank, world)
    a_ui, a_iu = shard.csrs(dev)
    kr, kc, kv = wl.knn_coo()
    mm_local = shard.mm_csr(kr, kc, kv, dev)
    ue = torch.from_numpy(wl.user_emb).to(dev)
    ie = torch.from_numpy(wl.item_emb[shard.local_items]).to(dev)
    px = sharded.PeerExchange.create(U, d, wl.n_layers, shard.n_local, dev, k=k)           # barriers inside the kernels
    px2 = sharded.PeerExchange.create(U, d, wl.n_layers, shard.n_local, dev, k=k)          # barriers as separate launches
    if px2 is not None:
        px2.sync_in_kernel = False
    have = torch.tensor([1.0 if (px is not None and px2 is not None) else 0.0], device=dev)
    dist.all_reduce(have, op=dist.ReduceOp.MIN)
    batches = []
    for lo in range(0, U, 1024):
        hi = min(U, lo + 1024)
        m = torch.from_numpy(wl.eval_mask(lo, hi)).to(dev)
        batches.append((torch.arange(lo, hi, device=dev), m, sharded.local_mask(shard, m)))
    results = {}
    for name in (["p2p", "p2p_launch_sync"] if have.item() else []) + ["nccl"]:
        with torch.no_grad():
            for rep in range(3):                                     # repeatedly: buffers and barrier flags are reused across calls
                if name.startswith("p2p"):
                    x = px if name == "p2p" else px2
                    u_g, i_g = sharded.propagate_mean_sharded_p2p(a_ui, a_iu, ue, ie, wl.n_layers, x, mm_local=mm_local)
                    out = [sharded.score_topk_sharded_p2p(shard, u_g, i_g, users, lm, k, x, bi * 1024) for bi, (users, _, lm) in enumerate(batches)]
                else:
                    u_g, i_g = sharded.propagate_mean_sharded(a_ui, a_iu, ue, ie, wl.n_layers)
                    i_g = sharded.mm_layer_sharded(shard, mm_local, ie, i_g)
                    out = []
                    for users, _, lm in batches:
                        v, i = sharded.score_topk_sharded(shard, u_g, i_g, users, lm, k, mask_is_local=True)
                        out.append((v, i, 0, users.numel()))
            torch.cuda.synchronize(); dist.barrier()
            results[name] = sharded.parity_vs_single_gpu(wl, shard, u_g, i_g, out, batches, kr, kc, kv, wl.n_layers, k, dev)
    ok = all(r["ok"] for r in results.values()) and len(results) > 0
    if rank == 0:
        print(json.dumps({"world": world, "peer_memory": bool(have.item()), "results": results, "ok": ok}))
    sys.stdout.flush()
    torch.cuda.synchronize(); dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
