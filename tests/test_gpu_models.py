"""Drop-in parity: the model classes of mmrec_b200 (same names / constructor / config keys as the reference's
src/models) on the GPU against the golden vectors recorded from the unmodified reference, and replayed training
trajectories.  Embeddings 1e-4 rel; losses 1e-5 rel; gradients 1e-4 rel; top-k near-tie rule; metrics equal."""
import os
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mmrec_oracle as O  # noqa: E402


def rel(a, b):
    a = a.detach().cpu().double() if torch.is_tensor(a) else torch.from_numpy(np.asarray(a)).double()
    b = b.detach().cpu().double() if torch.is_tensor(b) else torch.from_numpy(np.asarray(b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmrec_b200.utils import synth
    tmp = tempfile.mkdtemp(prefix="mmrec_gpu_")
    u, i, e, d, f = synth.SHAPES["tiny"]
    g = synth.make_graph(u, i, e, seed=0)
    v, t = synth.make_features(i, f, seed=1)
    synth.write_dataset(os.path.join(tmp, "data"), "tiny", g, v, t)
    return os.path.join(tmp, "data") + "/"


def build(model_name, data_path, overrides, gold=None):
    from mmrec_b200.utils.configurator import Config
    from mmrec_b200.utils.dataloader import EvalDataLoader, TrainDataLoader
    from mmrec_b200.utils.dataset import RecDataset
    from mmrec_b200.utils.utils import get_model, init_seed
    cfg = {"data_path": data_path, "eval_batch_size": 128, "train_batch_size": 512}
    cfg.update(overrides)
    config = Config(model_name, "tiny", cfg)
    for k in config["hyper_parameters"]:
        if isinstance(config[k], list):
            config[k] = config[k][0]
    ds = RecDataset(config)
    tr, va, te = ds.split()
    train = TrainDataLoader(config, tr, batch_size=config["train_batch_size"], shuffle=True)
    valid = EvalDataLoader(config, va, additional_dataset=tr, batch_size=config["eval_batch_size"])
    test = EvalDataLoader(config, te, additional_dataset=tr, batch_size=config["eval_batch_size"])
    init_seed(config["seed"])
    train.pretrain_setup()
    model = get_model(model_name)(config, train).to(config["device"])
    return config, train, valid, test, model


def load_gold_params(model, gold, prefix="param0."):
    sd = {k[len(prefix):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith(prefix)}
    model.load_state_dict(sd, strict=True)


def check_topk(idx, scores_ref, mask, k, scale):
    ref = torch.from_numpy(scores_ref).double().clone()
    ref[mask[0], mask[1]] = -1e10
    rv, ri = O.topk_tie_low_index(ref.numpy(), k)
    got = idx.cpu().numpy()
    for b in np.nonzero((got != ri).any(axis=1))[0]:
        cols = np.nonzero(got[b] != ri[b])[0]
        gap = np.abs(ref[b, got[b, cols]].numpy() - ref[b, ri[b, cols]].numpy()).max()
        assert gap < 1e-5 * scale, f"row {b}: top-k mismatch that is not a near tie (gap {gap})"


@pytest.mark.parametrize("name,file,over", [
    ("FREEDOM", "freedom_tiny.npz", {"n_ui_layers": 3}),
    ("BM3", "bm3_tiny.npz", {}),
    ("MGCN", "mgcn_tiny.npz", {}),
    ("LightGCN", "lightgcn_tiny.npz", {"n_layers": [3]}),
    ("LayerGCN", "layergcn_tiny.npz", {"dropout": [0.1]}),
])
def test_model_matches_reference(env, golden, name, file, over):
    gold = golden(file)
    config, train, valid, test, model = build(name, env, over)
    dev = config["device"]
    # (1) same seed + same construction order => the reference's initial weights, bit for bit
    for k, p in model.state_dict().items():
        assert np.array_equal(p.cpu().numpy(), gold["param0." + k]), f"initial {k} differs from the reference"
    assert [k for k, _ in model.named_parameters()] == list(gold["param_order"])
    # (2) forward
    model.eval()
    with torch.no_grad():
        if name in ("FREEDOM", "MGCN"):
            u, i = model.forward(model.norm_adj)
        elif name == "LayerGCN":
            model.forward_adj = model.norm_adj_matrix
            u, i = model.forward()
        else:
            u, i = model.forward()
    assert rel(u, gold["fwd_u"]) < 1e-5 and rel(i, gold["fwd_i"]) < 1e-5
    # (3) loss + gradients on the reference's recorded batch (graph pinned to the reference's draw)
    model.train()
    batch = torch.from_numpy(gold["batch"]).to(dev)
    if name in ("FREEDOM", "LayerGCN"):
        model.masked_adj = model.pruner.adj_from_keep(torch.from_numpy(gold["prune_keep_idx"]).to(dev))
    if name == "BM3":
        model.dropout = 0.0       # device RNG differs from the CPU stream: compare the dropout-free loss to the oracle
        p = {k[7:]: torch.from_numpy(gold[k]).requires_grad_(gold[k].dtype.kind == "f") for k in gold.files if k.startswith("param0.")}
        n = int(gold["n_users"]) + int(gold["n_items"])
        adj = torch.sparse_coo_tensor(torch.from_numpy(gold["norm_adj_idx"]), torch.from_numpy(gold["norm_adj_val"]), (n, n))
        ref_loss = O.bm3_loss(p, adj, torch.from_numpy(gold["batch"]), int(gold["cfg_n_layers"]), float(gold["cfg_reg_weight"]),
                              float(gold["cfg_cl_weight"]), 0.0)
        ref_loss.backward()
        ref_val = ref_loss.detach().numpy().reshape(-1)
        ref_grads = {k: v.grad.numpy() for k, v in p.items() if v.grad is not None and v.numel() <= 300 * 64}
    else:
        ref_val = gold["loss"]
        ref_grads = {k[5:]: gold[k] for k in gold.files if k.startswith("grad.")}
    model.zero_grad()
    loss = model.calculate_loss(batch)
    loss = sum(loss) if isinstance(loss, tuple) else loss
    loss.backward()
    np.testing.assert_allclose(loss.detach().cpu().numpy().reshape(-1), ref_val, rtol=2e-5)
    named = dict(model.named_parameters())
    assert ref_grads
    gmax = max(float(np.abs(g).max()) for g in ref_grads.values())
    for k, gref in ref_grads.items():
        # gradients that are pure cancellation noise (softmax-invariant biases, ~1e-11) carry no signal:
        # measure the error against the scale of the whole gradient, not of the vanishing entry
        err = (named[k].grad.detach().cpu().double() - torch.from_numpy(gref).double()).norm().item()
        assert err < 1e-4 * float(np.linalg.norm(gref)) + 1e-7 * gmax * np.sqrt(gref.size), f"grad {k}"
    # (4) full_sort_predict + trainer mask/top-k on the reference's first eval batch
    model.eval()
    with torch.no_grad():
        eb = [torch.from_numpy(gold["eval_users"]).to(dev), torch.from_numpy(gold["eval_mask"]).to(dev)]
        scores = model.full_sort_predict(eb)
        assert scores.shape == gold["scores"].shape and scores.dtype == torch.float32
        scale = float(np.abs(gold["scores"]).max())
        assert (scores.cpu() - torch.from_numpy(gold["scores"])).abs().max().item() < 2e-5 * scale
        s2 = model.full_sort_predict(eb)
        assert s2.data_ptr() != scores.data_ptr()            # fresh tensor each call: the trainer mutates it
        idx = model.full_sort_topk(eb, 50)
        check_topk(idx, gold["scores"], gold["eval_mask"], 50, scale)
    # (5) Trainer.evaluate: same metrics as the reference's trainer on valid and test
    from mmrec_b200.common.trainer import Trainer
    tr = Trainer(config, model)
    res = tr.evaluate(valid)
    got = np.array([res[k] for k in gold["metric_names"]])
    np.testing.assert_allclose(got, gold["metric_values"], atol=1e-4 + 1e-12)
    res_t = tr.evaluate(test)
    np.testing.assert_allclose(np.array([res_t[k] for k in gold["metric_names"]]), gold["test_metric_values"], atol=1e-4 + 1e-12)
    config["use_fused_topk"] = False                         # the reference's dense route gives the same metrics
    res_d = Trainer(config, model).evaluate(valid)
    assert res_d == res


@pytest.mark.parametrize("name,file,over", [
    ("LightGCN", "traj_lightgcn_tiny.npz", {"n_layers": [2], "reg_weight": [1e-4]}),
    ("FREEDOM", "traj_freedom_tiny.npz", {"dropout": [0.0], "reg_weight": [1e-3]}),
])
def test_training_trajectory_replay(env, golden, name, file, over):
    """Replay the batches the reference's dataloader produced through our model + Adam: per-batch losses,
    per-epoch Recall@20 (valid and test) and final embeddings must follow the reference's CPU run."""
    gold = golden(file)
    config, train, valid, test, model = build(name, env, over)
    from mmrec_b200.common.trainer import Trainer
    trainer = Trainer(config, model)
    dev = config["device"]
    batches = torch.from_numpy(gold["batches"])
    offs = np.concatenate([[0], np.cumsum(gold["batch_sizes"])])
    names = list(gold["metric_names"])
    b = 0
    for ep, nb in enumerate(gold["batches_per_epoch"]):
        model.pre_epoch_processing()
        model.train()
        for _ in range(int(nb)):
            inter = batches[:, offs[b]:offs[b + 1]].to(dev)
            trainer.optimizer.zero_grad()
            loss = model.calculate_loss(inter)
            np.testing.assert_allclose(loss.item(), gold["losses"][b], rtol=5e-5)
            loss.backward()
            trainer.optimizer.step()
            b += 1
        trainer.lr_scheduler.step()
        v = trainer.evaluate(valid)
        t = trainer.evaluate(test)
        np.testing.assert_allclose([v[k] for k in names], gold["valid"][ep], atol=2e-4)
        np.testing.assert_allclose([t[k] for k in names], gold["test"][ep], atol=2e-4)
        assert v["recall@20"] == pytest.approx(gold["valid"][ep][names.index("recall@20")], abs=1e-4)
    for k, p in model.state_dict().items():
        if "paramT." + k in gold.files:
            assert rel(p, gold["paramT." + k]) < 1e-4


def test_mmgcn_vs_oracle_restatement(env):
    """MMGCN has no reference pin (torch_geometric is absent): the PyG-free model class is checked against the oracle's
    torch-CPU restatement of the same file, forward + loss + gradients + scoring."""
    config, train, valid, test, model = build("MMGCN", env, {})
    dev = config["device"]
    assert model.concate and model.v_gcn.dim_latent == 256 and model.t_gcn.dim_latent is None
    p = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ei = model.edge_index.cpu()
    ide = model.id_embedding.detach().cpu()
    with torch.no_grad():
        out = model.forward()
        rv = O.mmgcn_gcn_forward(p, "v_gcn", ei, model.v_feat.cpu(), ide, model.v_gcn.preference.detach().cpu(), 256)
        rt = O.mmgcn_gcn_forward(p, "t_gcn", ei, model.t_feat.cpu(), ide, model.t_gcn.preference.detach().cpu(), None)
        ref = (rv + rt) / 2
    assert rel(out, ref) < 1e-4
    # mean aggregation alone, against the PyG semantics restated in the oracle
    from mmrec_b200 import ops
    x = torch.randn(ide.shape[0], 64)
    assert rel(ops.spmm(model.mean_adj, x.to(dev)), O.mmgcn_mean_aggregate(ei, x)) < 1e-5
    # loss + backward run, scoring uses the cached result
    batch = next(iter(train)).to(dev)
    model.train()
    loss = model.calculate_loss(batch)
    loss.backward()
    assert torch.isfinite(loss) and model.v_gcn.MLP.weight.grad is not None and model.t_gcn.conv_embed_1.weight.grad.abs().sum() > 0
    model.eval()
    with torch.no_grad():
        eb = next(iter(valid))
        s = model.full_sort_predict(eb)
        res = model.result.detach()
        assert rel(s, res[:model.n_users][eb[0]] @ res[model.n_users:].t()) < 1e-5
        idx = model.full_sort_topk(eb, 20)
        assert idx.shape == (eb[0].numel(), 20)


def test_quick_start_runs_end_to_end(env):
    """The whole drop-in flow (config -> data -> grid -> model -> trainer) on the GPU."""
    from mmrec_b200.utils.quick_start import quick_start
    results, best = quick_start("FREEDOM", "tiny", {"data_path": env, "epochs": 2, "dropout": [0.8], "reg_weight": [1e-3],
                                                    "eval_batch_size": 128, "train_batch_size": 512}, save_model=False)
    assert len(results) == 1
    assert 0.0 <= results[0][2]["recall@20"] <= 1.0 and results[0][1]["recall@20"] > 0.0


@pytest.mark.parametrize("name,file,attrs", [
    ("FREEDOM", "freedom_tiny.npz", ["mm_adj"]),
    ("MGCN", "mgcn_tiny.npz", ["image_original_adj", "text_original_adj"]),
])
def test_knn_item_graph_on_the_scoring_kernels(env, golden, name, file, attrs):
    """f4: the item-item graphs the models build at init (`graph._knn`: ops.score + ops.mask_topk) against the matrices the
    reference built from the same features (src/models/freedom.py:79-100, src/utils/utils.py:165-183): same coordinates,
    values to 1e-6."""
    gold = golden(file)
    config, train, valid, test, model = build(name, env, {})
    for attr in attrs:
        gi, gv = gold[attr + "_idx"], gold[attr + "_val"]
        n = getattr(model, attr).n_rows
        ref = torch.sparse_coo_tensor(torch.from_numpy(gi), torch.from_numpy(gv), (n, n)).coalesce()
        r, c, v = getattr(model, attr).coo()
        ours = torch.sparse_coo_tensor(torch.stack([r, c]).cpu(), v.cpu(), (n, n)).coalesce()
        assert torch.equal(ours.indices(), ref.indices()), f"{attr}: different neighbour sets"
        assert rel(ours.values(), ref.values()) < 1e-6


def test_mgcn_fused_inference_forward_equals_the_autograd_form(env):
    """a5b: `MGCN.forward` without autograd (gate / fuse kernels, SpMMs writing into the stacked tables) against the same
    forward under autograd (torch expressions of src/models/mgcn.py:153-201)."""
    config, train, valid, test, model = build("MGCN", env, {})
    model.eval()
    with torch.no_grad():
        u0, i0 = model.forward(model.norm_adj)
    u1, i1 = model.forward(model.norm_adj)                   # grad enabled: the torch route
    assert u1.requires_grad and not u0.requires_grad
    assert rel(u0, u1) < 2e-6 and rel(i0, i1) < 2e-6
