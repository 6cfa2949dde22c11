"""Host-side logic of the drop-in (no GPU): graph builders, config, dataset, dataloaders, evaluator -- against the
golden vectors recorded from the reference and against the oracle."""
import os
import random
import tempfile

import numpy as np
import pytest
import torch

from mmrec_b200 import graph
from mmrec_b200.utils import synth
from mmrec_b200.utils.configurator import Config
from mmrec_b200.utils.dataloader import EvalDataLoader, TrainDataLoader
from mmrec_b200.utils.dataset import RecDataset
from mmrec_b200.utils.topk_evaluator import TopKEvaluator, hit_matrix
from mmrec_b200.utils.utils import early_stopping, get_model
from oracle import mmrec_oracle as O


@pytest.fixture(scope="module")
def tiny_run():
    tmp = tempfile.mkdtemp(prefix="mmrec_host_")
    u, i, e, d, f = synth.SHAPES["tiny"]
    g = synth.make_graph(u, i, e, seed=0)
    v, t = synth.make_features(i, f, seed=1)
    synth.write_dataset(os.path.join(tmp, "data"), "tiny", g, v, t)
    cfg = Config("FREEDOM", "tiny", {"data_path": os.path.join(tmp, "data") + "/", "use_gpu": False,
                                     "eval_batch_size": 128, "train_batch_size": 512})
    ds = RecDataset(cfg)
    tr, va, te = ds.split()
    return cfg, g, tr, va, te


def test_synth_graph_matches_golden_graph(golden):
    gg = golden("tiny_graph.npz")
    g = synth.named("tiny")
    assert np.array_equal(g.user, gg["user"]) and np.array_equal(g.item, gg["item"]) and np.array_equal(g.label, gg["label"])
    tu, ti = g.train
    assert np.bincount(tu, minlength=g.n_users).min() >= 1 and np.bincount(ti, minlength=g.n_items).min() >= 1
    assert len(np.unique(g.user * g.n_items + g.item)) == len(g.user)


@pytest.mark.parametrize("f,attr", [("freedom_tiny.npz", "norm_adj"), ("lightgcn_tiny.npz", "norm_adj_matrix")])
def test_norm_adj_entries_bit_exact_vs_reference(golden, f, attr):
    g = golden(f)
    rows, cols, vals = graph.norm_adj_entries(g["inter_row"], g["inter_col"], int(g["n_users"]), int(g["n_items"]))
    assert np.array_equal(np.stack([rows, cols]), g[attr + "_idx"])
    assert np.array_equal(vals.view(np.uint32), g[attr + "_val"].view(np.uint32))


def test_mgcn_entries_bit_exact_vs_reference(golden):
    g = golden("mgcn_tiny.npz")
    rows, cols, vals = graph.mgcn_norm_adj_entries(g["inter_row"], g["inter_col"], int(g["n_users"]), int(g["n_items"]))
    assert np.array_equal(np.stack([rows, cols]), g["norm_adj_idx"])
    assert np.array_equal(vals.view(np.uint32), g["norm_adj_val"].view(np.uint32))


def test_config_layers_and_missing_keys(tiny_run):
    cfg = tiny_run[0]
    assert cfg["embedding_size"] == 64 and cfg["n_ui_layers"] == 2 and cfg["knn_k"] == 10
    assert cfg["reg_weight"] == [0.0, 1e-05, 1e-04, 1e-03]           # widened float syntax
    assert cfg["USER_ID_FIELD"] == "userID" and cfg["no_such_key"] is None
    assert cfg["hyper_parameters"] == ["seed", "dropout", "reg_weight"] or set(cfg["hyper_parameters"]) == {"seed", "dropout", "reg_weight"}
    assert cfg["valid_metric_bigger"] is True and cfg["device"].type == "cpu"
    cfg2 = Config("BM3", "tiny", {"n_layers": [2]})
    assert cfg2["use_neg_sampling"] is False and cfg2["n_layers"] == [2]


def test_dataset_split_and_cold_start_filter(tiny_run):
    cfg, g, tr, va, te = tiny_run
    assert len(tr) == int((g.label == 0).sum())
    train_users = set(tr.df["userID"].tolist())
    assert set(va.df["userID"]).issubset(train_users) and set(te.df["userID"]).issubset(train_users)
    assert tr.get_user_num() == g.n_users and tr.get_item_num() == g.n_items


def test_train_loader_batches_and_negatives(tiny_run):
    cfg, g, tr, va, te = tiny_run
    dl = TrainDataLoader(cfg, tr, batch_size=512, shuffle=True)
    dl.pretrain_setup()
    hist = set(zip(tr.df["userID"].tolist(), tr.df["itemID"].tolist()))
    seen = 0
    for batch in dl:
        assert batch.dtype == torch.int64 and batch.shape[0] == 3 and batch.shape[1] <= 512
        u, p, n = batch.numpy()
        assert all((a, b) in hist for a, b in zip(u.tolist(), p.tolist()))
        assert not any((a, b) in hist for a, b in zip(u.tolist(), n.tolist()))
        seen += batch.shape[1]
    assert seen == len(tr)
    m = dl.inter_matrix(form="coo")
    assert m.shape == (g.n_users, g.n_items) and m.nnz == len(tr)
    cfg_b = Config("BM3", "tiny", {"data_path": cfg["data_path"], "use_gpu": False})
    b = next(iter(TrainDataLoader(cfg_b, tr, batch_size=100)))
    assert b.shape == (2, 100)


def test_eval_loader_matches_reference_batches(tiny_run, golden):
    cfg, g, tr, va, te = tiny_run
    gold = golden("freedom_tiny.npz")
    dl = EvalDataLoader(cfg, va, additional_dataset=tr, batch_size=128)
    first = next(iter(dl))
    assert np.array_equal(first[0].numpy(), gold["eval_users"])
    # same mask entries per batch row (the reference keeps per-user order of appearance)
    assert np.array_equal(first[1].numpy(), gold["eval_mask"])
    dl.pr = 0
    n_users = 0
    for users, mask in dl:
        assert mask[0].max() < users.shape[0] and mask[0].min() >= 0
        n_users += users.shape[0]
    assert n_users == len(dl.get_eval_items()) == len(dl.get_eval_len_list())
    for a, b in zip(dl.get_eval_items()[:128], gold["eval_pos_items"]):
        assert sorted(a.tolist()) == sorted(b.tolist())


def test_evaluator_matches_reference_metrics(tiny_run, golden):
    cfg, g, tr, va, te = tiny_run
    gold = golden("lightgcn_tiny.npz")
    U, I = int(gold["n_users"]), int(gold["n_items"])
    adj = O.norm_adj_coo(gold["inter_row"], gold["inter_col"], U, I)
    ue, ie = torch.from_numpy(gold["param0.embedding_dict.user_emb"]), torch.from_numpy(gold["param0.embedding_dict.item_emb"])
    u, i = O.lightgcn_forward(adj, ue, ie, int(gold["cfg_n_layers"]))
    dl = EvalDataLoader(cfg, te, additional_dataset=tr, batch_size=128)
    lists = []
    for users, mask in dl:
        s = O.full_sort_scores(u, i, users)
        lists.append(O.mask_topk(s, mask, 50)[1])
    res = TopKEvaluator(cfg).evaluate(lists, dl)
    got = np.array([res[k] for k in gold["metric_names"]])
    np.testing.assert_allclose(got, gold["test_metric_values"], atol=1e-12)
    # vectorised hit matrix == the reference's membership loop
    topk = torch.cat(lists).numpy()
    slow = np.array([[x in set(m.tolist()) for x in n] for m, n in zip(dl.get_eval_items(), topk)])
    assert np.array_equal(hit_matrix(topk, dl.get_eval_items()), slow)


def test_early_stopping_and_plugin_loader():
    best, step, stop, upd = early_stopping(0.5, 0.4, 3, max_step=5, bigger=True)
    assert (best, step, stop, upd) == (0.5, 0, False, True)
    best, step, stop, upd = early_stopping(0.3, 0.5, 5, max_step=5, bigger=True)
    assert (best, step, stop, upd) == (0.5, 6, True, False)
    for name in ("FREEDOM", "BM3", "MGCN", "LightGCN", "LayerGCN", "MMGCN"):
        cls = get_model(name)
        assert cls.__name__ == name
        for meth in ("calculate_loss", "full_sort_predict", "forward", "pre_epoch_processing", "post_epoch_processing"):
            assert hasattr(cls, meth)


def test_models_fail_loudly_without_cuda(tiny_run):
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from mmrec_b200._lib import MMRecError
    cfg, g, tr, va, te = tiny_run
    dl = TrainDataLoader(cfg, tr, batch_size=512)
    for k in cfg["hyper_parameters"]:
        if isinstance(cfg[k], list):
            cfg[k] = cfg[k][0]
    with pytest.raises(MMRecError):
        get_model("FREEDOM")(cfg, dl)


def test_reference_arm_runs_on_rank_zero_only():
    """`bench.py --impl reference` under torchrun: every rank but 0 exits 0 without work and without output."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bm3_objective_is_the_references_op_sequence():
    """BM3's loss arithmetic after the hot-path kernels (`BM3._objective`) against the golden-pinned oracle `bm3_loss`
    on CPU tensors: same RNG draws, bit-identical loss and parameter gradients.  (n_layers = 0 makes the oracle's
    graph encoder the identity, so both sides see the same encoder outputs without a GPU.)"""
    import types
    import torch.nn.functional as F
    from oracle import mmrec_oracle as O
    from mmrec_b200.models.bm3 import BM3
    from mmrec_b200.common.loss import EmbLoss
    g = torch.Generator().manual_seed(5)
    U, I, d, Fv, Ft, B = 30, 25, 16, 40, 24, 64
    names = {"user_embedding.weight": (U, d), "item_id_embedding.weight": (I, d), "image_embedding.weight": (I, Fv),
             "text_embedding.weight": (I, Ft), "image_trs.weight": (d, Fv), "image_trs.bias": (d,), "text_trs.weight": (d, Ft),
             "text_trs.bias": (d,), "predictor.weight": (d, d), "predictor.bias": (d,)}
    base = {k: torch.randn(*shape, generator=g) * 0.3 for k, shape in names.items()}
    batch = torch.stack([torch.randint(0, U, (B,), generator=g), torch.randint(0, I, (B,), generator=g)])
    adj = torch.sparse_coo_tensor(torch.zeros(2, 0, dtype=torch.int64), torch.zeros(0), (U + I, U + I))
    results = []
    for side in ("oracle", "product"):
        p = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        torch.manual_seed(123)
        if side == "oracle":
            loss = O.bm3_loss(p, adj, batch, 0, 0.1, 2.0, 0.3)
        else:
            pred = types.SimpleNamespace(weight=p["predictor.weight"], bias=p["predictor.bias"])
            ns = types.SimpleNamespace(dropout=0.3, reg_weight=0.1, cl_weight=2.0, reg_loss=EmbLoss(), _apart=BM3._apart,
                                       predictor=lambda x: F.linear(x, pred.weight, pred.bias))
            u_all = p["user_embedding.weight"]
            i_all = p["item_id_embedding.weight"] + p["item_id_embedding.weight"]
            t_proj = F.linear(p["text_embedding.weight"], p["text_trs.weight"], p["text_trs.bias"])
            v_proj = F.linear(p["image_embedding.weight"], p["image_trs.weight"], p["image_trs.bias"])
            loss = BM3._objective(ns, u_all, i_all, t_proj, v_proj, batch[0], batch[1])
        loss.backward()
        results.append((loss.detach(), {k: v.grad for k, v in p.items()}))
    (l0, g0), (l1, g1) = results
    assert torch.equal(l0, l1)
    for k in names:
        assert torch.allclose(g0[k], g1[k], rtol=1e-6, atol=1e-9), k


def test_sharded_driver_only_uses_names_that_exist():
    """The multi-GPU driver and worker cannot run here (no GPUs); at least every `shard.<name>` / `px.<name>` / `ops.<name>` /
    `sharded.<name>` they spell must exist (a method lost in a refactor once broke N > 1 without any CPU test noticing)."""
    import inspect, re
    from mmrec_b200 import ops, sharded
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = inspect.getsource(sharded) + open(os.path.join(ROOT, "tests", "sharded_gpu_worker.py")).read() + open(os.path.join(ROOT, "bench.py")).read()
    for name in set(re.findall(r"\bshard\.([a-zA-Z_]\w*)", src)):
        assert hasattr(sharded.ItemShard, name) or name in ("rank", "world", "n_users", "n_items", "local_items", "n_local", "u", "i_local",
                                                            "val", "nnz"), f"ItemShard.{name} is used but not defined"
    px_attrs = set(re.findall(r"self\.([a-zA-Z_]\w*)", inspect.getsource(sharded.PeerExchange.__init__))) | set(dir(sharded.PeerExchange))
    for name in set(re.findall(r"\bpx\.([a-zA-Z_]\w*)", src)):
        assert name in px_attrs, f"PeerExchange.{name} is used but not defined"
    for name in set(re.findall(r"\bops\.([a-zA-Z_]\w*)", src)):
        assert hasattr(ops, name), f"ops.{name} is used but not defined"
    for name in set(re.findall(r"\bsharded\.([a-zA-Z_]\w*)", src)):
        assert hasattr(sharded, name), f"sharded.{name} is used but not defined"


def test_fused_adam_host_logic_with_cpu_stand_ins(monkeypatch):
    """`optim.FusedAdam` + `ops._ProjectFn.backward` host logic (bias corrections, factored table gradients, accumulation
    fallback, version bumps, state layout) with the five kernels replaced by torch-CPU restatements of their documented
    formulas (include/mmrec_b200.h, f1) -- against torch.optim.Adam on the same model.  The kernels themselves are checked on
    the GPU (tests/test_gpu_train.py); this keeps the glue honest on a CPU box."""
    from mmrec_b200 import ops
    from mmrec_b200.optim import FusedAdam

    def adam_el(p, g, m, v, b1, b2, eps, wd, step_size, bc2):
        g = g + wd * p if wd else g
        m.lerp_(g, 1 - b1); v.mul_(b2).addcmul_(g, g, value=1 - b2)
        p.add_(m / (v.sqrt() / bc2 + eps), alpha=step_size)
    monkeypatch.setattr(ops, "project_raw", lambda t, w, b, idx=None, l2=False: torch.nn.functional.linear(t if idx is None else t[idx], w, b))
    monkeypatch.setattr(ops, "linear_wgrad", lambda g, t, idx=None, want_bias=True: (g.t() @ (t if idx is None else t[idx]), g.sum(0) if want_bias else None))
    monkeypatch.setattr(ops, "index_sum_rows", lambda g, idx, n: torch.zeros(n, g.shape[1]).index_add_(0, idx, g))
    monkeypatch.setattr(ops, "linear_dgrad", lambda G, W: G @ W)
    monkeypatch.setattr(ops, "linear_dgrad_adam", lambda G, W, p, m, v, b1, b2, eps, wd, ss, bc2: adam_el(p, G @ W, m, v, b1, b2, eps, wd, ss, bc2))
    monkeypatch.setattr(ops, "adam_step", lambda entries, b1, b2, eps, wd: [adam_el(p, g, m, v, b1, b2, eps, wd, ss, bc2) for p, g, m, v, ss, bc2 in entries])

    def make():
        gen = torch.Generator().manual_seed(7)
        table = torch.nn.Parameter(torch.randn(60, 24, generator=gen))
        lin = torch.nn.Linear(24, 8)
        with torch.no_grad():
            lin.weight.copy_(0.3 * torch.randn(8, 24, generator=gen)); lin.bias.copy_(0.1 * torch.randn(8, generator=gen))
        emb = torch.nn.Parameter(torch.randn(60, 8, generator=gen))
        return table, lin, emb

    def loss(m, idx, ours):
        table, lin, emb = m
        proj = ops.project(table, lin.weight, lin.bias, idx=idx) if ours else torch.nn.functional.linear(table, lin.weight, lin.bias)[idx]
        return (proj * emb[idx]).sum(1).sigmoid().log().neg().mean()

    gen = torch.Generator().manual_seed(8)
    idxs = [torch.randint(0, 60, (90,), generator=gen) for _ in range(10)]
    for factored, accumulate, wd in ((True, False, 0.0), (True, True, 0.01), (False, False, 0.01)):
        a, b = make(), make()
        pa, pb = [a[0], *a[1].parameters(), a[2]], [b[0], *b[1].parameters(), b[2]]
        oa = FusedAdam(pa, lr=1e-2, weight_decay=wd, factored=factored)
        ob = torch.optim.Adam(pb, lr=1e-2, weight_decay=wd)
        sched_a = torch.optim.lr_scheduler.LambdaLR(oa, lambda e: 0.9 ** e)           # the trainer's scheduler works on the subclass
        sched_b = torch.optim.lr_scheduler.LambdaLR(ob, lambda e: 0.9 ** e)
        for s in range(5):
            oa.zero_grad(); ob.zero_grad()
            for r in range(2 if accumulate else 1):
                loss(a, idxs[2 * s + r], True).backward(); loss(b, idxs[2 * s + r], False).backward()
            if factored and not accumulate:
                assert a[0].grad is None and a[0]._mmrec_pending is not None
            v0 = a[0]._version
            oa.step(); ob.step(); sched_a.step(); sched_b.step()
            assert a[0]._version > v0 and a[0]._mmrec_pending is None
        for x, y in zip(pa, pb):
            np.testing.assert_allclose(x.detach().numpy(), y.detach().numpy(), rtol=2e-5, atol=2e-6)
            np.testing.assert_allclose(oa.state[x]["exp_avg_sq"].numpy(), ob.state[y]["exp_avg_sq"].numpy(), rtol=2e-5, atol=1e-9)
            assert float(oa.state[x]["step"]) == 5.0
        ob.load_state_dict(oa.state_dict())                                          # same layout: loads into torch's Adam
        oa.release()
        assert all(getattr(p, "_mmrec_defer", None) is None for p in pa)


def test_unique_sorted_is_np_unique():
    rng = np.random.default_rng(3)
    for n, hi in ((0, 10), (1, 1), (1000, 50), (200000, 10**12)):
        x = rng.integers(0, hi, n)
        assert np.array_equal(graph.unique_sorted(x), np.unique(x))
    # the config-5-shaped builder path: symmetric keys of a bipartite graph, entries identical to the np.unique formulation
    r, c = rng.integers(0, 5000, 60000), rng.integers(0, 3000, 60000)
    rows, cols, vals = graph.norm_adj_entries(r, c, 5000, 3000)
    n = 8000
    k = np.unique(np.concatenate([r * n + (c + 5000), (c + 5000) * n + r]))
    assert np.array_equal(rows, k // n) and np.array_equal(cols, k % n) and vals.dtype == np.float32


def test_every_ops_and_graph_name_spelled_in_the_package_exists():
    """No GPU here, so most of `mmrec_b200` cannot be executed: at least every `ops.<name>` / `graph.<name>` / `_lib.<name>` the
    package, bench.py, the tools and the entry points spell must exist, and every C entry point `ops` calls must be bound."""
    import re
    from mmrec_b200 import _lib, graph as G, ops
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(dp, f) for top in ("mmrec_b200", "tools") for dp, _, fs in os.walk(os.path.join(root, top)) for f in fs if f.endswith(".py")]
    files += [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")]
    for path in files:
        src = open(path).read()
        for name in set(re.findall(r"(?<![\w.])ops\.([a-zA-Z_]\w*)", src)):
            assert hasattr(ops, name), f"{path}: ops.{name} does not exist"
        for name in set(re.findall(r"(?<![\w.])graph\.([a-zA-Z_]\w*)", src)):
            assert hasattr(G, name), f"{path}: graph.{name} does not exist"
        for name in set(re.findall(r"(?<![\w.])_lib\.([a-zA-Z_]\w*)", src)):
            assert hasattr(_lib, name), f"{path}: _lib.{name} does not exist"
        for name in set(re.findall(r"\.(mmrec_[a-z0-9_]+)\(", src)):
            assert name in _lib.PROTOTYPES, f"{path}: C entry point {name} is called but not bound in _lib.PROTOTYPES"
