"""Generate golden input/output vectors by running the LIVE reference (CPU).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so the vectors it produces are
committed as small .npz fixtures next to this script.  The reference ships no
tests for this path; these fixtures are what pins the oracle (oracle/fq_oracle.py).

Import recipe: bypass `src/quantization/__init__.py`
(which drags in Lightning) by pre-registering empty package modules, and stub the
stray `from matplotlib import scale` of gdnsq.py:1.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("MHAQ_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name, sub in [("src", "src"), ("src.quantization", "src/quantization"),
                      ("src.quantization.gdnsq", "src/quantization/gdnsq"),
                      ("src.quantization.gdnsq.layers", "src/quantization/gdnsq/layers"),
                      ("src.aux", "src/aux")]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = m
    mpl = types.ModuleType("matplotlib")
    mpl.scale = None
    sys.modules["matplotlib"] = mpl
    from src.quantization.gdnsq import gdnsq  # noqa
    from src.quantization.gdnsq.layers import gdnsq_act, gdnsq_conv2d, gdnsq_linear  # noqa
    from src.quantization.gdnsq.gdnsq_utils import QNMethod
    from src.aux.types import QScheme
    return gdnsq, gdnsq_act, gdnsq_conv2d, gdnsq_linear, QNMethod, QScheme


class FixedNoise:
    """Replace torch.randint_like by a seeded {0,1} generator and record the draw."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.last = None
        self._orig = torch.randint_like

    def __enter__(self):
        def fake(t, high, **kw):
            assert high == 2
            self.last = torch.randint(0, 2, t.shape, generator=self.gen).to(t.dtype)
            return self.last.clone()
        torch.randint_like = fake
        return self

    def __exit__(self, *a):
        torch.randint_like = self._orig


def t2n(t):
    return None if t is None else t.detach().cpu().numpy()


ONLY = set(sys.argv[1:])     # optional: names of the fixtures to (re)write; default all


def save(name, **arrs):
    if ONLY and name not in ONLY:
        return
    arrs = {k: v for k, v in arrs.items() if v is not None}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(f"  {name}: " + ", ".join(f"{k}{tuple(np.shape(v))}" for k, v in arrs.items()))


def main():
    torch.manual_seed(1234)
    torch.set_num_threads(1)
    gdnsq, gact, gconv, glin, QNMethod, QScheme = import_reference()

    # ---------------- activations: NoisyAct (always QNSTE, gdnsq_quant.py:508-511) -------------
    act_cases = [
        # name, signed, log_s, log_q, act_b, x scale
        ("act_signed_4b", True, -2.0, 2.0, -2.0, 1.0),
        ("act_unsigned_4b", False, -2.25, 1.75, 0.0, 1.0),     # non power-of-two scale
        ("act_signed_8b_odd", True, -5.3, 2.7, -3.1, 1.5),
        ("act_signed_1b", True, 1.0, 2.0, -2.0, 1.0),
        ("act_zero_bits", True, 1.0, 1.0, -1.0, 1.0),          # q == s -> lo == hi
        ("act_neg_bits", True, 1.5, 1.0, -1.0, 1.0),           # q < s  -> lo > hi
    ]
    for i, (name, signed, ls, lq, b, xs) in enumerate(act_cases):
        g = torch.Generator().manual_seed(100 + i)
        x = torch.randn(2, 3, 20, 17, generator=g) * xs
        if not signed:
            x = torch.relu(x)
        # exact ties at .5 and values on the clip bounds
        x.view(-1)[:8] = torch.tensor([b, b + 2 ** ls * 0.5, b + 2 ** ls * 1.5, b + 2 ** ls * 2.5,
                                       b + 2 ** lq - 2 ** ls, b + 2 ** lq, b - 1.0, 0.0])
        go = torch.randn(x.shape, generator=g)
        act = gact.NoisyAct(signed=signed)
        with torch.no_grad():
            act.log_act_s.fill_(ls)
            act.log_act_q.fill_(lq)
            act.act_b.fill_(b)
        act.train()
        xr = x.clone().requires_grad_(True)
        with FixedNoise(7 + i) as fn:
            y = act(xr)
            y.backward(go)
            bits = fn.last
        # codes through the two-call API (eval mode: exercises the asserts) and bw
        act.eval()
        with torch.no_grad():
            act(x)
            codes = act.Q.quantize(x)
            bw = act.bw
        save(name, x=t2n(x), go=t2n(go), log_act_s=np.float32(ls), log_act_q=np.float32(lq),
             act_b=np.float32(b), signed=np.bool_(signed), noise_bits=t2n(bits),
             y=t2n(y), gx=t2n(xr.grad), g_log_act_s=t2n(act.log_act_s.grad),
             g_log_act_q=t2n(act.log_act_q.grad),
             g_act_b=t2n(act.act_b.grad) if act.act_b.grad is not None else None,
             codes=t2n(codes), bw=t2n(bw))

    # ---------------- weights: NoisyConv2d weight path ----------------------------------------
    w_cases = [
        ("w_pc_ste", "PER_CHANNEL", "STE", (8, 4, 3, 3), 4, False),
        ("w_pc_lsq", "PER_CHANNEL", "LSQ", (8, 4, 3, 3), 3, False),
        ("w_pc_aewgs", "PER_CHANNEL", "AEWGS", (8, 4, 3, 3), 2, False),
        ("w_pc_aewgs_1b", "PER_CHANNEL", "AEWGS", (6, 5, 3, 3), 1, False),   # rows of 45: ragged
        ("w_pt_ste", "PER_TENSOR", "STE", (8, 4, 3, 3), 4, False),
        ("w_pt_lsq", "PER_TENSOR", "LSQ", (5, 3, 3, 3), 2, False),
        ("w_pc_ste_bias", "PER_CHANNEL", "STE", (8, 4, 3, 3), 4, True),
        # quantized bias under AEWGS: value (O,), scale (O,) -> reduce_to_shape has no singleton
        # dim, torch.mean(dim=()) reduces over everything: per-TENSOR statistics for the bias
        ("w_pc_aewgs_bias", "PER_CHANNEL", "AEWGS", (8, 4, 3, 3), 2, True),
        # per-tensor AEWGS: scale (1,) -> statistics over dim 0 only
        ("w_pt_aewgs", "PER_TENSOR", "AEWGS", (6, 4, 3, 3), 3, False),
    ]
    for i, (name, scheme, method, shape, bits_w, qbias) in enumerate(w_cases):
        g = torch.Generator().manual_seed(200 + i)
        O, I, kh, kw = shape
        conv = gconv.NoisyConv2d(I, O, (kh, kw), bias=True, qscheme=QScheme[scheme],
                                 quant_bias=qbias, qnmethod=QNMethod[method])
        with torch.no_grad():
            conv.weight.copy_(torch.randn(shape, generator=g) * 0.1)
            conv.bias.copy_(torch.randn(O, generator=g) * 0.05)
            # duplicate the row minimum in row 0 to exercise amin's tie split
            conv.weight[0, 1, 1, 1] = conv.weight[0].min()
            if scheme == "PER_CHANNEL":
                mx = conv.weight.amax((1, 2, 3), keepdim=True)
                mn = conv.weight.amin((1, 2, 3), keepdim=True)
                conv.log_wght_s.copy_(torch.log2((mx - mn) / (2 ** bits_w - 1)))
            else:
                conv.log_wght_s.fill_(float(torch.log2((conv.weight.max() - conv.weight.min())
                                                       / (2 ** bits_w - 1))))
        captured = {}

        def fake_conv(inp, w, b, _c=captured):
            _c["w"], _c["b"] = w, b
            return w.sum() * 0

        conv._conv_forward = fake_conv
        conv.train()
        go = torch.randn(shape, generator=g)
        gob = torch.randn(O, generator=g)
        with FixedNoise(50 + i) as fn:
            conv(torch.zeros(1, I, 8, 8))
            wq, bq = captured["w"], captured["b"]
            loss = (wq * go).sum()
            if qbias:
                loss = loss + (bq * gob).sum()
            # the weight-path draw happens first in backward order? record all draws
            draws = []
            orig = torch.randint_like

            def rec(t, high, **kw):
                r = orig(t, high, **kw)
                draws.append(r.clone())
                return r
            torch.randint_like = rec
            loss.backward()
            torch.randint_like = orig
        noise = {tuple(d.shape): d for d in draws}
        save(name, weight=t2n(conv.weight), bias=t2n(conv.bias), log_wght_s=t2n(conv.log_wght_s),
             go=t2n(go), go_bias=t2n(gob) if qbias else None,
             noise_bits=t2n(noise.get(tuple(shape))),
             noise_bits_bias=t2n(noise.get((O,))) if qbias else None,
             wq=t2n(wq), bq=t2n(bq) if qbias else None,
             g_weight=t2n(conv.weight.grad), g_log_wght_s=t2n(conv.log_wght_s.grad),
             g_bias=t2n(conv.bias.grad) if qbias else None,
             per_channel=np.bool_(scheme == "PER_CHANNEL"), method=np.str_(method))

    # ---------------- NoisyLinear, per-tensor (the only usable scheme) -------
    g = torch.Generator().manual_seed(250)
    lin = glin.NoisyLinear(20, 12, bias=True, qscheme=QScheme.PER_TENSOR, qnmethod=QNMethod.LSQ)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(12, 20, generator=g) * 0.3)
        lin.bias.copy_(torch.randn(12, generator=g) * 0.1)
        lin.log_wght_s.fill_(float(torch.log2((lin.weight.max() - lin.weight.min()) / (2 ** 4 - 1))))
    lin.train()
    xl = torch.randn(5, 20, generator=g).requires_grad_(True)
    gol = torch.randn(5, 12, generator=g)
    yl = lin(xl)
    yl.backward(gol)
    save("lin_pt_lsq", weight=t2n(lin.weight), bias=t2n(lin.bias), log_wght_s=t2n(lin.log_wght_s),
         x=t2n(xl), go=t2n(gol), y=t2n(yl), gx=t2n(xl.grad), g_weight=t2n(lin.weight.grad),
         g_bias=t2n(lin.bias.grad), g_log_wght_s=t2n(lin.log_wght_s.grad))

    # ---------------- raw Quantizer with explicit tensors (two-call API) ----------------------
    g = torch.Generator().manual_seed(300)
    x = torch.randn(4, 6, 5, 5, generator=g)
    scale = (torch.rand(4, 1, 1, 1, generator=g) * 0.2 + 0.05).requires_grad_(True)
    zp = (-torch.rand(4, 1, 1, 1, generator=g)).requires_grad_(True)

    class M:  # stand-in for the owning module (only .training is read)
        training = True
    Q = gdnsq.Quantizer(M(), scale, zp, -float("inf"), float("inf"), qnmethod=QNMethod.LSQ)
    xr = x.clone().requires_grad_(True)
    codes = Q.quantize(xr)
    gcodes = torch.randn(x.shape, generator=g)
    codes.backward(gcodes)
    save("quantizer_codes_lsq", x=t2n(x), scale=t2n(scale), zp=t2n(zp), gcodes=t2n(gcodes),
         codes=t2n(codes), gx=t2n(xr.grad), g_scale=t2n(scale.grad), g_zp=t2n(zp.grad))


def step_fixture():
    """One whole loss evaluation + backward of a two-layer quantized model through the reference's
    layers, ModelHelper.get_model_values (utils/model_helper.py:11-76) and PotentialLoss
    (gdnsq_loss.py:6-88): pins rows (f)-1 of SURVEY.md §8 — the range term
    log2(max - min + 2^log_wght_s), the constraint loss, and how autograd accumulates the
    parameter gradients across the three users of every scale."""
    from collections import OrderedDict
    gdnsq, gact, gconv, glin, QNMethod, QScheme = import_reference()
    for name, sub in [("src.quantization.gdnsq.utils", "src/quantization/gdnsq/utils")]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = m
    from src.quantization.gdnsq.utils.model_helper import ModelHelper
    from src.quantization.gdnsq.gdnsq_loss import PotentialLoss
    nn = torch.nn
    g = torch.Generator().manual_seed(400)

    def block(cin, cout, signed, bias):
        return nn.Sequential(OrderedDict([
            ("activations_quantizer", gact.NoisyAct(signed=signed)),
            ("0", gconv.NoisyConv2d(cin, cout, 3, padding=1, bias=bias, qscheme=QScheme.PER_CHANNEL,
                                    qnmethod=QNMethod.LSQ))]))
    model = nn.Sequential(OrderedDict([("c1", block(3, 8, True, True)), ("relu", nn.ReLU()),
                                       ("c2", block(8, 4, False, False))]))
    # integer-valued log parameters: exp2 is exact on every device
    act_params = {"c1": (-3.0, 2.0, -2.0), "c2": (-2.0, 1.0, 0.0)}     # log_act_s, log_act_q, act_b
    with torch.no_grad():
        for k, (ls, lq, b) in act_params.items():
            a = getattr(model, k).activations_quantizer
            a.log_act_s.fill_(ls); a.log_act_q.fill_(lq); a.act_b.fill_(b)
        for k in ("c1", "c2"):
            conv = getattr(model, k)._modules["0"]
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.2)
            conv.weight[0, 0, 0, 0] = conv.weight[0].min()         # tie at a row minimum
            conv.weight[1, 0, 1, 1] = conv.weight[1].max()         # tie at a row maximum
            if conv.bias is not None:
                conv.bias.copy_(torch.randn(conv.bias.shape, generator=g) * 0.05)
            rng = conv.weight.amax((1, 2, 3), keepdim=True) - conv.weight.amin((1, 2, 3), keepdim=True)
            conv.log_wght_s.copy_(torch.floor(torch.log2(rng / (2 ** 6 - 1))))   # ~6-7 bits, integer log
    x = torch.randn(2, 3, 10, 10, generator=g)
    target = torch.randn(2, 4, 10, 10, generator=g)
    crit = PotentialLoss(nn.MSELoss(), p=1, a=4, w=4)
    crit.t, crit.loss_sum, crit.cnt = 0.5, 0.8, 2          # calib_mul = 0.4: constraint term is live
    model.train(); crit.train()
    draws = []
    orig = torch.randint_like
    gen = torch.Generator().manual_seed(77)

    def rec(t, high, **kw):
        r = torch.randint(0, 2, t.shape, generator=gen).to(t.dtype)
        draws.append(r.clone())
        return r
    torch.randint_like = rec
    try:
        out = model(x)
        vals = ModelHelper.get_model_values(model, QScheme.PER_CHANNEL)
        loss = crit((out, *vals), target)
        loss.backward()
    finally:
        torch.randint_like = orig
    noise = {tuple(d.shape): d for d in draws}
    arrs = dict(x=t2n(x), target=t2n(target), out=t2n(out), loss=t2n(loss), wloss=t2n(crit.wloss),
                aloss=t2n(crit.aloss), log_act_s=t2n(vals[0]), log_act_q=t2n(vals[1]),
                log_wght_s=t2n(vals[2]), log_w_n_b=t2n(vals[3]),
                noise_bits_c1=t2n(noise[(2, 3, 10, 10)]), noise_bits_c2=t2n(noise[(2, 8, 10, 10)]))
    for n_, p_ in model.named_parameters():
        arrs["p:" + n_] = t2n(p_)
        if p_.grad is not None:
            arrs["g:" + n_] = t2n(p_.grad)
    save("step_pc_lsq", **arrs)


if __name__ == "__main__":
    main()
    if not ONLY or "step_pc_lsq" in ONLY:
        step_fixture()
