"""Shared helpers: golden fixtures, oracle drivers, comparison rules."""
import glob
import math
import os

import numpy as np
import torch

from oracle import fq_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name, device="cpu"):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        if a.dtype.kind in "US":
            out[k] = str(a)
        elif a.dtype == np.bool_:
            out[k] = bool(a)
        else:
            out[k] = torch.from_numpy(np.array(a)).to(device)
    return out


def bits_to_noise(bits):
    """randint_like(v,2).sub_(0.5) given the recorded {0,1} draw."""
    return None if bits is None else bits.float() - 0.5


from oracle.checks import assert_bit_exact, assert_close_rel, assert_param_grad  # noqa: E402,F401


def _pin_to_cpu_value(t, log_t):
    """exp2 is parameter preparation OUTSIDE the kernels and is done by torch on either
    device; CUDA's exp2f and the CPU's vectorised exp2 may differ in the last bit, so when
    comparing a CUDA run with CPU-generated fixtures give the CUDA graph the CPU value."""
    if t.is_cuda:
        t.data.copy_(torch.exp2(log_t.detach().cpu()))


# ---------------------------------------------------------------------------
# Drivers: run one golden case through a backend.  `fq` is a function with the
# signature of oracle.fq_oracle.fake_quant / mhaq_b200.ops.fake_quant.
# ---------------------------------------------------------------------------
def run_act_case(c, fq, device="cpu"):
    """NoisyAct.forward training step (gdnsq_act.py:42-55) with explicit noise."""
    x = c["x"].to(device).clone().requires_grad_(True)
    log_s = torch.tensor([float(c["log_act_s"])], device=device, requires_grad=True)
    log_q = torch.tensor([float(c["log_act_q"])], device=device, requires_grad=True)
    act_b = torch.tensor([float(c["act_b"])], device=device, requires_grad=c["signed"])
    s = torch.exp2(log_s)
    q = torch.exp2(log_q)
    _pin_to_cpu_value(s, log_s)
    _pin_to_cpu_value(q, log_q)
    y = fq(x, s, act_b, act_b, act_b + q - s, method="STE", noise=bits_to_noise(c["noise_bits"]).to(device))
    y.backward(c["go"].to(device))
    return dict(y=y, gx=x.grad, g_log_act_s=log_s.grad, g_log_act_q=log_q.grad,
                g_act_b=act_b.grad if c["signed"] else None)


def run_weight_case(c, fq, device="cpu"):
    """NoisyConv2d weight (and bias) path (gdnsq_conv2d.py:72-98) with explicit noise."""
    w = c["weight"].to(device).clone().requires_grad_(True)
    log_s = c["log_wght_s"].to(device).clone().requires_grad_(True)
    s = torch.exp2(log_s)
    _pin_to_cpu_value(s, log_s)
    zp = w.amin((1, 2, 3), keepdim=True) if c["per_channel"] else w.amin()
    noise = bits_to_noise(c.get("noise_bits"))
    wq = fq(w, s, zp, -math.inf, math.inf, method=c["method"],
            noise=None if noise is None else noise.to(device))
    loss = (wq * c["go"].to(device)).sum()
    out = {}
    if "bq" in c:
        b = c["bias"].to(device).clone().requires_grad_(True)
        nb = bits_to_noise(c["noise_bits_bias"]).to(device)
        bq = fq(b, s.ravel(), zp.ravel(), -math.inf, math.inf, method=c["method"], noise=nb)
        loss = loss + (bq * c["go_bias"].to(device)).sum()
        out["bq"] = bq
    loss.backward()
    out.update(wq=wq, g_weight=w.grad, g_log_wght_s=log_s.grad)
    if "bq" in c:
        out["g_bias"] = b.grad
    return out


def exact_act_param_grads(c):
    """fp64 sums of the reference's fp32 per-element terms, chained to the log-params in
    fp64 (gdnsq_act.py:42-48): the value both fp32 implementations approximate."""
    import math as _m
    x = c["x"]
    sv = torch.exp2(torch.tensor([float(c["log_act_s"])]))
    qv = torch.exp2(torch.tensor([float(c["log_act_q"])]))
    bv = torch.tensor([float(c["act_b"])])
    full = lambda t: t.expand(x.shape).contiguous().requires_grad_(True)
    s_, z_, l_, h_ = full(sv), full(bv), full(bv), full((bv + qv) - sv)
    y = O.fake_quant(x, s_, z_, l_, h_, "STE", bits_to_noise(c["noise_bits"]))
    y.backward(c["go"])
    gS, gZ, gL, gH = (t.grad.double().sum() for t in (s_, z_, l_, h_))
    ln2 = _m.log(2.0)
    return dict(g_log_act_s=((gS - gH) * sv.double() * ln2).reshape(1),
                g_log_act_q=(gH * qv.double() * ln2).reshape(1),
                g_act_b=(gZ + gL + gH).reshape(1))

