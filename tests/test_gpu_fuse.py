"""a5b (csrc/fuse.cu): MGCN's row-wise fusion kernels against the torch expressions of src/models/mgcn.py:153-154,187-201."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmrec_b200 import _lib
    _lib.require_device()
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("n,d,with_mul", [(23000, 64, True), (1, 64, True), (37, 32, False), (1001, 128, True), (4, 64, False)])
def test_gate_rows(dev, n, d, with_mul):
    from mmrec_b200 import ops
    torch.manual_seed(n + d)
    gate = nn.Sequential(nn.Linear(d, d), nn.Sigmoid()).to(dev)
    x, mul = torch.randn(n, d, device=dev), torch.randn(n, d, device=dev)
    with torch.no_grad():
        want = torch.sigmoid(x.double() @ gate[0].weight.double().t() + gate[0].bias.double())
        want = want * mul.double() if with_mul else want
        got = ops.gate_rows(x, gate[0].weight, gate[0].bias, mul=mul if with_mul else None)
    assert got.shape == (n, d) and rel(got, want) < 2e-6


@pytest.mark.parametrize("n,d", [(63000, 64), (5, 64), (333, 32), (1000, 128), (1, 64)])
def test_mgcn_fuse(dev, n, d):
    from mmrec_b200 import ops
    torch.manual_seed(n + d)
    q = nn.Sequential(nn.Linear(d, d), nn.Tanh(), nn.Linear(d, 1, bias=False)).to(dev)
    gi, gt = nn.Sequential(nn.Linear(d, d), nn.Sigmoid()).to(dev), nn.Sequential(nn.Linear(d, d), nn.Sigmoid()).to(dev)
    img, txt, content = (torch.randn(n, d, device=dev) for _ in range(3))
    with torch.no_grad():
        def lin(m, x):
            return x @ m.weight.double().t() + (0 if m.bias is None else m.bias.double())
        i64, t64, c64 = img.double(), txt.double(), content.double()                         # mgcn.py:187-201 in fp64
        att = torch.cat([lin(q[2], torch.tanh(lin(q[0], i64))), lin(q[2], torch.tanh(lin(q[0], t64)))], dim=-1)
        w = torch.softmax(att, dim=-1)
        common = w[:, 0].unsqueeze(1) * i64 + w[:, 1].unsqueeze(1) * t64
        side = (torch.sigmoid(lin(gi[0], c64)) * (i64 - common) + torch.sigmoid(lin(gt[0], c64)) * (t64 - common) + common) / 3
        want = c64 + side
        got, got_side = ops.mgcn_fuse(img, txt, content, q[0].weight, q[0].bias, q[2].weight, gi[0].weight, gi[0].bias, gt[0].weight,
                                      gt[0].bias, want_side=True)
    assert rel(got, want) < 2e-6 and rel(got_side, side) < 5e-6


def test_fuse_rejects_unsupported_width(dev):
    from mmrec_b200 import ops
    from mmrec_b200._lib import MMRecError
    with pytest.raises(MMRecError):
        ops.gate_rows(torch.zeros(4, 48, device=dev), torch.zeros(48, 48, device=dev), None)
