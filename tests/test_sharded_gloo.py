"""Multi-process logic of the item-sharded path on CPU (gloo, world_size 2 and 3): partition, per-layer
all-reduce, top-k all-gather + merge.  The local kernels are injected with oracle-backed CPU stand-ins -- the
orchestration and the collectives under test are the product's (mmrec_b200/sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mmrec_oracle as O


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class CpuMat:
    def __init__(self, r, c, v, n_rows, n_cols):
        self.t = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])), torch.from_numpy(v), (n_rows, n_cols))
        self.n_rows = n_rows


def cpu_spmm(A, X, acc_in=None, acc_div=1.0, want_y=True):
    y = torch.sparse.mm(A.t, X)
    acc = None
    if acc_in is not None:
        acc_in.copy_((acc_in + y) / acc_div)
        acc = acc_in
    return (y if want_y else None), acc


def cpu_score_topk(user_e, item_e, users, mask, k):
    s = O.full_sort_scores(user_e, item_e, users)
    if mask is not None and mask.numel():
        s[mask[0], mask[1]] = -1e10
    v, i = O.topk_tie_low_index(s.numpy(), min(k, s.shape[1]))
    return torch.from_numpy(v.copy()), torch.from_numpy(i.copy())


def cpu_merge(vals, idxs):
    parts, B, k = vals.shape
    v = vals.permute(1, 0, 2).reshape(B, parts * k).numpy()
    i = idxs.permute(1, 0, 2).reshape(B, parts * k).numpy()
    order = np.lexsort((i, -v.astype(np.float64)), axis=1)[:, :k]
    return torch.from_numpy(np.take_along_axis(v, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1))


def worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mmrec_b200 import sharded
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "freedom_tiny.npz"), allow_pickle=True)
        U, I = int(g["n_users"]), int(g["n_items"])
        ue = torch.from_numpy(g["param0.user_embedding.weight"]); ie = torch.from_numpy(g["param0.item_id_embedding.weight"])
        sh = sharded.ItemShard(g["inter_row"], g["inter_col"], U, I, rank, world)
        assert np.array_equal(sh.to_global(np.arange(sh.n_local)), sh.local_items)
        a_ui = CpuMat(sh.u, sh.i_local, sh.val, U, sh.n_local)
        a_iu = CpuMat(sh.i_local, sh.u, sh.val, sh.n_local, U)
        L = 3
        u_g, i_loc = sharded.propagate_mean_sharded(a_ui, a_iu, ue.clone(), ie[sh.local_items].clone(), L, spmm=cpu_spmm)
        ref = O.propagate_mean(O.norm_adj_coo(g["inter_row"], g["inter_col"], U, I), torch.cat([ue, ie]), L)
        assert ((u_g - ref[:U]).norm() / ref[:U].norm()).item() < 1e-5
        assert ((i_loc - ref[U:][sh.local_items]).norm() / ref[U:][sh.local_items].norm()).item() < 1e-5
        # the item-item layer of the sharded FREEDOM (freedom.py:166-167,178): all-gather of the layer-0 item rows (rank-major),
        # local SpMM with this rank's rows of mm_adj -> i_g + mm_adj @ E_I
        if I % world == 0:
            mm_idx, mm_val = g["mm_adj_idx"], g["mm_adj_val"]
            r, c, v = sh.mm_coo(mm_idx[0], mm_idx[1], mm_val)
            mm_local = CpuMat(r, c, v, sh.n_local, world * sh.n_local)
            i_full = sharded.mm_layer_sharded(sh, mm_local, ie[sh.local_items].clone(), i_loc.clone(), spmm=cpu_spmm)
            mm = torch.sparse_coo_tensor(torch.from_numpy(mm_idx), torch.from_numpy(mm_val), (I, I))
            ref_i = ref[U:] + torch.sparse.mm(mm, ie)
            assert ((i_full - ref_i[sh.local_items]).norm() / ref_i[sh.local_items].norm()).item() < 1e-5
        # every shard's edge count adds up to the graph
        n = torch.tensor([sh.nnz]); dist.all_reduce(n)
        assert int(n.item()) == len(np.unique(g["inter_row"] * I + g["inter_col"]))
        # global top-k == single-process top-k on the same embeddings
        users = torch.from_numpy(g["eval_users"]); mask = torch.from_numpy(g["eval_mask"])
        v, idx = sharded.score_topk_sharded(sh, u_g, i_loc, users, mask, 20, score_topk=cpu_score_topk, merge=cpu_merge)
        full_i = torch.empty(I, ie.shape[1])
        parts = [torch.empty(len(range(r, I, world)), ie.shape[1]) for r in range(world)]
        dist.all_gather(parts, i_loc)
        for r in range(world):
            full_i[r::world] = parts[r]
        rv, ri = cpu_score_topk(u_g, full_i, users, mask, 20)
        assert torch.equal(idx, ri) and torch.allclose(v, rv, rtol=0, atol=0)
        # the mask pre-filtered to the shard (what the bench / a per-rank loader passes) gives the same lists
        lm = sharded.local_mask(sh, mask)
        assert lm.shape[1] == int(((mask[1] % world) == rank).sum()) and torch.all(lm[1] < sh.n_local)
        v2, idx2 = sharded.score_topk_sharded(sh, u_g, i_loc, users, lm, 20, score_topk=cpu_score_topk, merge=cpu_merge,
                                              mask_is_local=True)
        assert torch.equal(idx2, idx) and torch.equal(v2, v)
        assert sharded.local_mask(sh, None) is None
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_item_sharded_path_under_gloo(tmp_path, world):
    port = free_port()
    mp.spawn(worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
