"""Pins oracle/convlstm_oracle.py against the UNMODIFIED reference classes (build container only)."""
import pytest
import torch

from oracle import convlstm_oracle as O
from oracle.reference_loader import load_reference

pytestmark = pytest.mark.reference


def _ref_model(params, cin, hid, cout):
    _, Net, _ = load_reference()
    net = Net(cin, hid, cout)
    net.load_state_dict(params)
    return net


@pytest.mark.parametrize("hid,hw,tin,tout,B", [(32, (16, 16), 4, 4, 2), (16, (7, 9), 3, 2, 1)])
def test_forward_bit_exact(hid, hw, tin, tout, B):
    torch.manual_seed(1)
    p = O.init_params(12, hid, 12, seed=0)
    x = torch.randn(B, tin, 12, *hw)
    net = _ref_model(p, 12, hid, 12)
    with torch.no_grad():
        y_ref = net(x, tout)
        y, sv = O.rollout_forward(x, p, tout)
    assert y.shape == y_ref.shape
    assert torch.equal(y, y_ref)


def test_backward_matches_reference_autograd():
    torch.manual_seed(2)
    hid, tin, tout, B = 16, 3, 4, 2
    p = O.init_params(12, hid, 5, seed=3, cell_weight_scale=3.0)
    x = torch.randn(B, tin, 12, 12, 10)
    tgt = torch.rand(B, tout, 5, 12, 10)
    net = _ref_model(p, 12, hid, 5)
    y_ref = net(x, tout)
    loss_ref = torch.nn.MSELoss()(y_ref.permute(0, 2, 1, 3, 4), tgt)
    loss_ref.backward()
    y, sv = O.rollout_forward(x, p, tout)
    loss, dy = O.mse_loss_and_grad(y, tgt)
    assert abs(loss.item() - loss_ref.item()) < 1e-7
    g = O.rollout_backward(dy, sv, p)
    ref_g = {k: v.grad for k, v in net.named_parameters()}
    for k in p:
        assert O.rel_l2(g[k], ref_g[k]) < 2e-5, (k, O.rel_l2(g[k], ref_g[k]))


def test_cell_matches_reference():
    Cell, _, _ = load_reference()
    torch.manual_seed(0)
    cell = Cell(12, 8, (3, 5), True)
    x, h, c = torch.randn(2, 12, 7, 9), torch.randn(2, 8, 7, 9), torch.randn(2, 8, 7, 9)
    with torch.no_grad():
        h_ref, c_ref = cell(x, [h, c])
        hn, cn, _ = O.cell_forward(x, h, c, cell.conv.weight, cell.conv.bias)
    assert torch.equal(hn, h_ref) and torch.equal(cn, c_ref)


def test_known_answer_gate_order():
    """SURVEY Appendix C probes: block 3 is tanh g, block 2 is o."""
    hid = 4
    w = torch.zeros(4 * hid, 3 + hid, 3, 3)
    b = torch.zeros(4 * hid)
    b[0:hid] = 10.0
    b[3 * hid :] = 10.0
    z = torch.zeros(1, hid, 5, 5)
    hn, cn, _ = O.cell_forward(torch.zeros(1, 3, 5, 5), z, z, w, b)
    assert torch.allclose(cn, torch.ones_like(cn), atol=1e-3)
    b = torch.zeros(4 * hid)
    b[2 * hid : 3 * hid] = 10.0
    hn, cn, _ = O.cell_forward(torch.zeros(1, 3, 5, 5), z, torch.ones_like(z), w, b)
    assert torch.allclose(hn, torch.full_like(hn, 0.4621), atol=1e-3)
