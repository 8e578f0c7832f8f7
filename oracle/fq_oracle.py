"""CPU oracle for the fake-quantization hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or as
the timed CPU baseline — never as part of the product path.  The product
(``mhaq_b200``) has no CPU fallback and never imports ``oracle``.

What it restates (paths relative to the reference checkout, aifoundry-org/MHAQ):

* ``quantize`` / ``dequantize``  — ``Quantizer.quantize`` / ``.dequantize``,
  src/quantization/gdnsq/gdnsq.py:189-229
* ``_RoundNoise``                — ``QNoise`` and its four estimator subclasses
  ``QNSTE`` / ``QNLSQ`` / ``QNEWGS`` / ``QNAEWGS``, gdnsq.py:11-147, and
  ``reduce_to_shape``, gdnsq.py:150-152
* ``act_fake_quant``             — ``NoisyAct.forward``, layers/gdnsq_act.py:39-55
* ``weight_fake_quant``          — the weight path of ``NoisyConv2d.forward`` /
  ``NoisyLinear.forward``, layers/gdnsq_conv2d.py:71-98, layers/gdnsq_linear.py:61-76

It is a restatement, not a copy: one functional autograd ``Function`` covers the
estimators, and the same ATen ops are issued in the same order so that (a) fp32
results are bit-identical to the reference's on the same device, and (b) timing it
on host cores is representative of the reference's CPU path.

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against outputs of the reference itself:
``tests/golden/make_golden.py`` imports the live reference in the build container
and stores seeded input/output vectors under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this module against them (bit-exact
forward / input-gradient, ≤1e-6 relative parameter gradients).

Deliberate difference: the reference's ``QNEWGS.backward`` raises ``AttributeError``
on every call (typo ``ctx.need_input_grad``, gdnsq.py:102); the oracle implements
the evident intent (same scale gradient as the STE estimator).  Parity for EWGS
is therefore against this oracle only.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist

STE, EWGS, AEWGS, LSQ = 0, 1, 2, 3
METHODS = {"STE": STE, "EWGS": EWGS, "AEWGS": AEWGS, "LSQ": LSQ}


def _mid(method) -> int:
    if isinstance(method, str):
        return METHODS[method]
    if isinstance(method, int):
        return method
    return int(method.value)


def reduce_to_shape(t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """Mean over every dim where `like` has size 1 (gdnsq.py:150-152)."""
    dims = tuple(i for i, n in enumerate(like.shape) if n == 1)
    return torch.mean(t, dim=dims, keepdim=True)


def draw_noise(v: torch.Tensor) -> torch.Tensor:
    """r in {-0.5, +0.5}, i.i.d. (gdnsq.py:54)."""
    return torch.randint_like(v, 2).sub_(0.5)


class _RoundNoise(torch.autograd.Function):
    """forward: round(v) - v  (gdnsq.py:13-16); backward: the chosen estimator."""

    @staticmethod
    def forward(ctx, v, scale, method: int, noise: Optional[torch.Tensor]):
        ctx.save_for_backward(v, scale)
        ctx.method = method
        ctx.noise = noise
        return torch.round(v) - v

    @staticmethod
    def backward(ctx, g):
        v, scale = ctx.saved_tensors
        method = ctx.method
        gv = gs = None
        need_v, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1]

        def gdnsq_scale_grad():
            # gdnsq.py:54-55 — stochastic scale gradient of arXiv:2508.14004
            r = ctx.noise if ctx.noise is not None else draw_noise(v)
            return (3.0 ** -0.5) * g * r

        if method == STE:
            if need_v:
                gv = g * 0                                     # gdnsq.py:50
            if need_s:
                gs = gdnsq_scale_grad()
        elif method == LSQ:
            if need_v:
                gv = g * 0                                     # gdnsq.py:78
            if need_s:
                gs = g * (torch.round(v) - v)                  # gdnsq.py:81-82
        elif method == EWGS:
            if need_v:
                e = torch.round(v) - v
                gv = -torch.abs(g) * e * 1e-2                  # gdnsq.py:96-100
            if need_s:
                gs = gdnsq_scale_grad()                        # intent of gdnsq.py:102-105
        elif method == AEWGS:
            if need_v:
                e = torch.round(v) - v                         # gdnsq.py:118
                num_full = g.sign() * e
                e2_full = e.square()
                num = reduce_to_shape(num_full, scale).detach()
                e2 = reduce_to_shape(e2_full, scale).detach()
                me = reduce_to_shape(e, scale).detach()
                if dist.is_available() and dist.is_initialized():
                    dist.all_reduce(num, op=dist.ReduceOp.AVG)   # gdnsq.py:126-129
                    dist.all_reduce(e2, op=dist.ReduceOp.AVG)
                    dist.all_reduce(me, op=dist.ReduceOp.AVG)
                den = (e2 - me.square()).clamp_min(1e-3)
                delta = num / den
                g_scale = (1.0 * delta * num_full).clamp_max(1 - 0.01)
                gv = -g * g_scale                               # gdnsq.py:141
            if need_s:
                gs = gdnsq_scale_grad()
        else:
            raise AttributeError(f"Unknown method {method}!")
        return gv, gs, None, None


def quantize(value, scale, zero_point, min_val, max_val, method="STE", noise=None):
    """Integer-valued codes, differentiable (gdnsq.py:189-219, training branch)."""
    value = torch.clamp(value, min=min_val, max=max_val)       # :197
    value = value - zero_point                                 # :199
    value = value / scale                                      # :204
    value = value + _RoundNoise.apply(value, scale, _mid(method), noise)   # :206-208
    return value


def check_codes(codes, scale, zero_point, min_val, max_val):
    """The eval-mode assertions of gdnsq.py:211-217."""
    if torch.any(codes < torch.floor((min_val - zero_point) / scale)):
        raise AssertionError("Not all elements in the tensor above min val")
    if torch.any(codes > torch.ceil((max_val - zero_point) / scale)):
        raise AssertionError("Not all elements in the tensor below max val")
    if not torch.all((codes == codes.floor()) | (codes == codes.ceil())):
        raise AssertionError("Not all elements in the tensor have integer values.")


def dequantize(codes, scale, zero_point):
    return codes * scale + zero_point                          # gdnsq.py:229


def fake_quant(value, scale, zero_point, min_val=-math.inf, max_val=math.inf, method="STE",
               noise=None):
    return dequantize(quantize(value, scale, zero_point, min_val, max_val, method, noise),
                      scale, zero_point)


def act_fake_quant(x, log_act_s, log_act_q, act_b, noise=None, method="STE"):
    """NoisyAct.forward in training mode (gdnsq_act.py:42-55)."""
    s = torch.exp2(log_act_s)
    q = torch.exp2(log_act_q)
    zero_point = act_b
    min_val = act_b
    max_val = act_b + q - s
    return fake_quant(x, s, zero_point, min_val, max_val, method, noise)


def act_bit_width(codes):
    """NoisyAct eval-mode `bw` (gdnsq_act.py:51-54)."""
    mm = codes.aminmax()
    return torch.log2(mm.max - mm.min + 1)


def weight_fake_quant(weight, log_wght_s, per_channel: bool, method="STE", noise=None):
    """Weight path of NoisyConv2d.forward (gdnsq_conv2d.py:72-84, 98)."""
    s = torch.exp2(log_wght_s)
    if per_channel:
        zp = weight.amin(tuple(range(1, weight.dim())), keepdim=True)
    else:
        zp = weight.amin()
    return fake_quant(weight, s, zp, -math.inf, math.inf, method, noise)


def bias_fake_quant(bias, weight, log_wght_s, method="STE", noise=None):
    """quant_bias branch (per-channel only): reuses the weight scale / row-min,
    ravelled (gdnsq_conv2d.py:86-94)."""
    s = torch.exp2(log_wght_s)
    zp = weight.amin(tuple(range(1, weight.dim())), keepdim=True)
    return fake_quant(bias, s.ravel(), zp.ravel(), -math.inf, math.inf, method, noise)


# ---------------------------------------------------------------------------
# Philox4x32-10 noise stream of the CUDA kernels (numpy restatement of
# mhaq_b200/csrc/fq_common.cuh: noise_block / noise_nibble)
# ---------------------------------------------------------------------------
def philox_noise(n_rows: int, n_inner: int, seed: int, offset: int):
    """r[row, p] in {-0.5,+0.5}: one Philox block per (row, super-tile of 16384, thread of 128);
    bit 4*it+k belongs to element k of the float4 the thread owns in iteration it."""
    import numpy as np

    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = 0x9E3779B9, 0xBB67AE85
    mask = np.uint64(0xFFFFFFFF)
    supers = (n_inner + 16383) // 16384
    p = np.arange(n_inner, dtype=np.uint64)
    T = p >> np.uint64(14)
    within = p & np.uint64(16383)
    it = (within >> np.uint64(9)).astype(np.int64)
    tid = (within >> np.uint64(2)) & np.uint64(127)
    k = (within & np.uint64(3)).astype(np.int64)
    out = np.empty((n_rows, n_inner), dtype=np.float32)
    for row in range(n_rows):
        pos = (np.uint64(row) * np.uint64(supers) + T) * np.uint64(128) + tid
        c0 = pos & mask
        c1 = (pos >> np.uint64(32)) & mask
        c2 = np.full_like(c0, offset & 0xFFFFFFFF)
        c3 = np.full_like(c0, (offset >> 32) & 0xFFFFFFFF)
        k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
        for _ in range(10):
            p0 = M0 * c0
            p1 = M1 * c2
            n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & mask
            n1 = p1 & mask
            n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & mask
            n3 = p0 & mask
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = (k0 + W0) & 0xFFFFFFFF
            k1 = (k1 + W1) & 0xFFFFFFFF
        words = np.stack([c0, c1, c2, c3], axis=0)            # [4, n_inner]
        w = words[it >> 3, np.arange(n_inner)]
        bit = (w >> ((it & 7) * 4 + k).astype(np.uint64)) & np.uint64(1)
        out[row] = bit.astype(np.float32) - 0.5
    return out
