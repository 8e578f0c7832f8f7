"""CPU: the oracle restatement vs vectors produced by the live reference
(tests/golden/make_golden.py).  Forward / input gradients bit-exact; parameter
gradients within 1e-6 relative (they are fp32 reductions whose order is the same
ATen kernels here and there, so in practice they are identical too)."""
import math

import pytest
import torch

from oracle import fq_oracle as O
from tests import helpers as H


@pytest.mark.parametrize("name", H.golden_names("act_"))
def test_act_golden(name):
    c = H.load_golden(name)
    r = H.run_act_case(c, O.fake_quant)
    H.assert_bit_exact(r["y"], c["y"], "y")
    H.assert_bit_exact(r["gx"], c["gx"], "gx")
    for k in ("g_log_act_s", "g_log_act_q", "g_act_b"):
        if k in c:
            H.assert_close_rel(r[k], c[k], 1e-6, k, abs_floor=1e-7)
    # two-call API + eval extras
    s = torch.exp2(torch.tensor([float(c["log_act_s"])]))
    q = torch.exp2(torch.tensor([float(c["log_act_q"])]))
    b = torch.tensor([float(c["act_b"])])
    codes = O.quantize(c["x"], s, b, b, b + q - s)
    H.assert_bit_exact(codes, c["codes"], "codes")
    O.check_codes(codes, s, b, b, b + q - s)
    H.assert_bit_exact(O.act_bit_width(codes), c["bw"], "bw")


@pytest.mark.parametrize("name", H.golden_names("w_"))
def test_weight_golden(name):
    c = H.load_golden(name)
    r = H.run_weight_case(c, O.fake_quant)
    H.assert_bit_exact(r["wq"], c["wq"], "wq")
    H.assert_close_rel(r["g_weight"], c["g_weight"], 1e-6, "g_weight", abs_floor=1e-7)
    H.assert_close_rel(r["g_log_wght_s"], c["g_log_wght_s"], 1e-6, "g_log_wght_s", abs_floor=1e-7)
    if "bq" in c:
        H.assert_bit_exact(r["bq"], c["bq"], "bq")
        H.assert_close_rel(r["g_bias"], c["g_bias"], 1e-6, "g_bias", abs_floor=1e-7)


def test_quantizer_codes_golden():
    c = H.load_golden("quantizer_codes_lsq")
    x = c["x"].clone().requires_grad_(True)
    scale = c["scale"].clone().requires_grad_(True)
    zp = c["zp"].clone().requires_grad_(True)
    codes = O.quantize(x, scale, zp, -math.inf, math.inf, "LSQ")
    codes.backward(c["gcodes"])
    H.assert_bit_exact(codes, c["codes"], "codes")
    H.assert_bit_exact(x.grad, c["gx"], "gx")
    H.assert_close_rel(scale.grad, c["g_scale"], 1e-6, "g_scale", abs_floor=1e-7)
    H.assert_close_rel(zp.grad, c["g_zp"], 1e-6, "g_zp", abs_floor=1e-7)


def test_ewgs_intent_runs():
    """The reference's QNEWGS cannot run (gdnsq.py:102); the oracle's restatement of its
    intent must at least produce the documented closed form."""
    torch.manual_seed(0)
    v = (torch.randn(64) * 3).requires_grad_(True)
    s = torch.tensor([0.5], requires_grad=True)
    r = torch.randint(0, 2, (64,)).float() - 0.5
    out = O._RoundNoise.apply(v, s, O.EWGS, r)
    g = torch.randn(64)
    out.backward(g)
    e = torch.round(v.detach()) - v.detach()
    # total grad to v = grad of (round(v) - v) path only here: -|g| * e * 0.01
    assert torch.equal(v.grad, -torch.abs(g) * e * 1e-2)
    torch.testing.assert_close(s.grad, ((3.0 ** -0.5) * g * r).sum().reshape(1))


def test_whole_step_golden_oracle_and_loss_mirror():
    """tests/golden/step_pc_lsq.npz — layers + ModelHelper.get_model_values + PotentialLoss of the
    live reference on a two-convolution model: the oracle's restatement of the layers and the
    product's PotentialLoss mirror (plain torch, device-agnostic) reproduce the loss and every
    parameter gradient, including how autograd sums the three uses of each weight scale."""
    from torch import nn
    from mhaq_b200.quantization.gdnsq.gdnsq_loss import PotentialLoss
    c = H.load_golden("step_pc_lsq")
    # like the generator: the CPU convolution's weight-gradient reduction order depends on the
    # thread count, and d/d log_wght_s amplifies that last-bit noise (nearly cancelling sums)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        _whole_step_checks(c, PotentialLoss, nn)
    finally:
        torch.set_num_threads(threads)


def _whole_step_checks(c, PotentialLoss, nn):
    P, out, vals = H.run_step_case_oracle(c)
    H.assert_bit_exact(out, c["out"], "model output")
    for v, k in zip(vals, ("log_act_s", "log_act_q", "log_wght_s", "log_w_n_b")):
        H.assert_bit_exact(v, c[k], k)
    crit = PotentialLoss(nn.MSELoss(), p=1, a=4, w=4)
    loss = H.step_case_loss(crit, out, vals, c["target"])
    H.assert_close_rel(loss, c["loss"], 1e-6, "loss")
    H.assert_close_rel(crit.wloss, c["wloss"], 1e-6, "wloss")
    H.assert_close_rel(crit.aloss, c["aloss"], 1e-6, "aloss")
    assert float(crit.cnt) == 3 and abs(float(crit.loss_sum) - (0.8 + crit.rloss.item())) < 1e-6
    loss.backward()
    n = 0
    for k, g in c.items():
        if k.startswith("g:"):
            H.assert_close_rel(P[k[2:]].grad, g, 1e-6, k, abs_floor=1e-7)
            n += 1
    assert n == 10
    assert P["c1.0.log_b_s"].grad is None and P["c2.activations_quantizer.act_b"].grad is None


def test_noisy_linear_golden():
    """NoisyLinear.forward, per-tensor LSQ (gdnsq_linear.py:61-78), run by the live reference."""
    import torch.nn.functional as F
    c = H.load_golden("lin_pt_lsq")
    w = c["weight"].clone().requires_grad_(True)
    b = c["bias"].clone().requires_grad_(True)
    ls = c["log_wght_s"].clone().requires_grad_(True)
    x = c["x"].clone().requires_grad_(True)
    y = F.linear(x, O.weight_fake_quant(w, ls, False, "LSQ"), b)
    y.backward(c["go"])
    H.assert_bit_exact(y, c["y"], "y")
    H.assert_bit_exact(x.grad, c["gx"], "gx")
    H.assert_close_rel(w.grad, c["g_weight"], 1e-6, "g_weight", abs_floor=1e-7)
    H.assert_close_rel(b.grad, c["g_bias"], 1e-6, "g_bias", abs_floor=1e-7)
    H.assert_close_rel(ls.grad, c["g_log_wght_s"], 1e-6, "g_log_wght_s", abs_floor=1e-7)
