"""Seeded differential fuzz: the CUDA path (through mhaq_b200.fake_quant -> the C ABI) against the
LIVE reference (oracle/_ref) on the same GPU, over randomly drawn shapes, layouts, parameter
shapes, estimators, bit widths, scale magnitudes, clamp ranges and awkward values.

The parity tests elsewhere pin named configurations; this one walks the space between them:
ragged / tiny / misaligned sizes, per-tensor and per-channel (dim 0) parameters, clamp ranges that
are finite, one-sided or absent, channels_last and sliced (non-contiguous) inputs, inputs sitting
exactly on rounding ties and on the clamp bounds, zeros, signed zeros and huge magnitudes.

Bars: y and the STE / LSQ input gradient bit-exact; parameter gradients by
oracle/checks.py:assert_param_grad (1e-5 of the reference's fp32 value, or as close as the
reference itself to the fp64 sum of its own fp32 terms).
"""
import math
import os
import random
import types

import pytest
import torch

from oracle import checks as C
from oracle import fq_oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.gpu

REL = 1e-5
N_CASES = int(os.environ.get("MHAQ_FUZZ_CASES", "96"))     # more seeds: MHAQ_FUZZ_CASES=1000 pytest ...


@pytest.fixture(scope="module")
def ref():
    if not ref_loader.available():
        pytest.skip("oracle/_ref not staged")
    return ref_loader.load_ops()


@pytest.fixture(scope="module")
def fq():
    import mhaq_b200
    return mhaq_b200


def _draw_shape(rng):
    kind = rng.choice(["vec", "mat", "conv_w", "act", "tiny", "ragged"])
    if kind == "vec":
        return (rng.choice([1, 3, 4, 511, 2044, 2048, 4099, 70001, 1 << 18]),)
    if kind == "mat":
        return (rng.choice([1, 2, 7, 64, 300]), rng.choice([1, 5, 64, 577, 4096, 9001]))
    if kind == "conv_w":
        return (rng.choice([1, 8, 48, 130]), rng.choice([1, 3, 16, 50]), rng.choice([1, 3]), rng.choice([1, 3]))
    if kind == "act":
        return (rng.choice([1, 2, 5]), rng.choice([3, 16, 33]), rng.choice([7, 14, 17]), rng.choice([7, 14, 17]))
    if kind == "tiny":
        return tuple(rng.choice([1, 2, 3]) for _ in range(rng.choice([1, 2, 3, 4])))
    return (rng.choice([2, 3, 5]), rng.choice([4096 + 1000, 8192 + 3, 12289]))


def _layout(rng, x):
    """Return a tensor with the same values in a randomly chosen memory layout."""
    how = rng.choice(["contig", "contig", "channels_last", "sliced", "transposed"])
    if how == "channels_last" and x.dim() == 4:
        return x.contiguous(memory_format=torch.channels_last)
    if how == "sliced" and x.dim() >= 1:
        pad = torch.zeros(x.shape[:-1] + (x.shape[-1] + 3,), device=x.device)
        pad[..., 1:1 + x.shape[-1]] = x
        return pad[..., 1:1 + x.shape[-1]]
    if how == "transposed" and x.dim() == 2:
        return x.t().contiguous().t()
    return x.contiguous()


def _case(seed):
    rng = random.Random(seed)
    g = torch.Generator(device="cuda").manual_seed(seed)
    shape = _draw_shape(rng)
    method = rng.choice(["STE", "STE", "LSQ"])
    bits = rng.choice([1, 2, 3, 4, 5, 8])
    per_channel = len(shape) >= 2 and rng.random() < 0.5
    pshape = ((shape[0],) + (1,) * (len(shape) - 1)) if per_channel else (1,)
    mag = 2.0 ** rng.choice([-9, -3, 0, 0, 2, 6])
    x = torch.randn(shape, device="cuda", generator=g) * mag
    go = torch.randn(shape, device="cuda", generator=g) * 2.0 ** rng.choice([-20, -6, 0, 0, 5])
    lo = (-2.0 * mag * (0.5 + torch.rand(pshape, device="cuda", generator=g)))
    width = 4.0 * mag * (0.5 + torch.rand(pshape, device="cuda", generator=g))
    scale = width / (2 ** bits - 1 if bits > 1 else 1)
    if rng.random() < 0.3:                       # a power-of-two scale: many exact ties
        scale = torch.exp2(torch.round(torch.log2(scale)))
    hi = lo + scale * (2 ** bits - 1)
    zp = lo.clone()
    clamp = rng.choice(["both", "both", "none", "lo_only", "hi_only"])
    awkward = rng.random() < 0.6
    if awkward and x.numel() >= 8:
        flat = x.reshape(-1)
        k = max(1, flat.numel() // 16)
        idx = torch.randperm(flat.numel(), device="cuda", generator=g)
        s0 = scale.reshape(-1)[0]
        z0 = zp.reshape(-1)[0]
        ties = z0 + s0 * (torch.randint(0, 2 ** bits, (k,), device="cuda", generator=g).float() + 0.5)
        flat[idx[:k]] = ties                                         # exact rounding ties (channel 0's grid)
        flat[idx[k:2 * k]] = 0.0
        flat[idx[2 * k:2 * k + 1]] = -0.0
        flat[idx[2 * k + 1:2 * k + 2]] = float(lo.reshape(-1)[0])    # exactly on the bounds
        flat[idx[2 * k + 2:2 * k + 3]] = float(hi.reshape(-1)[0])
        # far outside: beyond 2^80 (the kernels' exact-division range guard) when the clamp catches it
        flat[idx[2 * k + 3:2 * k + 4]] = 3.0e30 if clamp == "both" else 1.0e4 * mag
        go.reshape(-1)[idx[:1]] = 0.0
    x = _layout(rng, x)
    go = _layout(rng, go)
    return dict(shape=shape, method=method, bits=bits, per_channel=per_channel, clamp=clamp,
                x=x, go=go, scale=scale, zp=zp, lo=lo, hi=hi, seed=seed)


def _kernel_noise(fq, x, scale, seed, offset):
    """The noise the kernels draw for `x`, as a tensor indexed like x.  The stream is a function of
    the STORAGE position inside the [rows][inner] view: a dense channels_last
    tensor is walked as it lies in memory (NHWC), anything else in row-major order."""
    if x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last):
        nhwc = x.permute(0, 2, 3, 1)                               # contiguous view of the storage
        sc = scale.reshape(-1, 1, 1, 1) if scale.numel() > 1 else None
        return fq.philox_noise(nhwc, sc, seed=seed, offset=offset).permute(0, 3, 1, 2)
    return fq.philox_noise(x.contiguous(), scale if scale.numel() > 1 else None, seed=seed, offset=offset)


@pytest.mark.parametrize("seed", range(N_CASES))
def test_fuzz_against_the_live_reference(fq, ref, seed):
    c = _case(1000 + seed)
    x, go, scale, zp = c["x"], c["go"], c["scale"], c["zp"]
    use_lo = c["clamp"] in ("both", "lo_only")
    use_hi = c["clamp"] in ("both", "hi_only")
    method = c["method"]
    r = None if method == "LSQ" else _kernel_noise(fq, x, scale, 11, seed)
    orig_randint_like = torch.randint_like

    def leaves():
        mk = lambda t: t.detach().clone().requires_grad_(True)
        return mk(scale), mk(zp), (mk(c["lo"]) if use_lo else None), (mk(c["hi"]) if use_hi else None)

    # ---- live reference
    xr = x.detach().clone().requires_grad_(True)
    s_r, z_r, l_r, h_r = leaves()
    # (torch.clamp takes two tensors or two numbers: a one-sided range is a tensor of +-inf there)
    inf = torch.full_like(c["lo"], math.inf)
    none = not (use_lo or use_hi)
    Q = ref.Quantizer(types.SimpleNamespace(training=True), s_r, z_r,
                      l_r if use_lo else (-math.inf if none else -inf),
                      h_r if use_hi else (math.inf if none else inf), qnmethod=ref.QNMethod[method])
    if r is not None:
        torch.randint_like = lambda t, high, **kw: (r.to(t.dtype) + 0.5)
    try:
        y_r = Q.dequantize(Q.quantize(xr))
        y_r.backward(go)
    finally:
        torch.randint_like = orig_randint_like
    # ---- CUDA path
    xo = x.detach().clone().requires_grad_(True)
    s_o, z_o, l_o, h_o = leaves()
    y_o = fq.fake_quant(xo, s_o, z_o, l_o if use_lo else -math.inf, h_o if use_hi else math.inf,
                        method=method, philox=(11, seed))
    y_o.backward(go)
    tag = f"seed {seed}: {c['shape']} {method} {c['bits']}b {'per-channel' if c['per_channel'] else 'per-tensor'} clamp={c['clamp']}"
    C.assert_bit_exact(y_o, y_r, "y " + tag)
    C.assert_bit_exact(xo.grad, xr.grad, "gx " + tag)
    # ---- parameter gradients
    n_per = x.numel() // scale.numel()
    ex = C.exact_param_grads(O.fake_quant, x.detach().contiguous(), go.contiguous(), scale, zp,
                             c["lo"] if use_lo else (None if none else -inf),
                             c["hi"] if use_hi else (None if none else inf), method,
                             None if r is None else r.contiguous())
    # fp32 rounding of the per-element terms: go*code and g*(v/s) are formed at the magnitude
    # |go| * |v|, the estimator's term (3^-1/2 * g * r, or g * e for LSQ) at up to 0.5 * |go| * s
    xc = x.detach()
    if use_lo:
        xc = torch.maximum(xc, c["lo"])
    if use_hi:
        xc = torch.minimum(xc, c["hi"])
    vmax = float(((xc - zp) / scale).abs().max())
    gmax = float(go.abs().max()) + 1e-30
    floor = 2e-7 * math.sqrt(n_per) * gmax * (vmax + 4 + 0.5 * float(scale.abs().max()))
    C.assert_param_grad(s_o.grad, s_r.grad, ex[0], REL, "g_scale " + tag, floor)
    C.assert_param_grad(z_o.grad, z_r.grad, ex[1], REL, "g_zp " + tag, floor)
    if use_lo:
        C.assert_param_grad(l_o.grad, l_r.grad, ex[2], REL, "g_lo " + tag, floor)
    if use_hi:
        C.assert_param_grad(h_o.grad, h_r.grad, ex[3], REL, "g_hi " + tag, floor)


@pytest.mark.parametrize("seed", range(32))
def test_fuzz_per_channel_weights_all_estimators(fq, ref, seed):
    """Weight-style tensors (per-channel, no clamp, calibration-formula scale) at random [C, inner]
    shapes and bit widths, every estimator the reference can run — AEWGS included, whose input
    gradient depends on per-channel fp32 means (1e-5 relative instead of bit-exact)."""
    from tests.test_gpu_live_reference import _check_per_channel, _per_channel_case
    rng = random.Random(7000 + seed)
    C_ = rng.choice([1, 2, 3, 16, 48, 64, 130, 512])
    inner = rng.choice([1, 2, 9, 27, 144, 576, 1152, 4095, 4096, 4608, 8192 + 5, 20000])
    if C_ * inner < 2:
        inner = 9
    bits = rng.choice([1, 2, 3, 4, 8])
    method = rng.choice(["STE", "LSQ", "AEWGS", "AEWGS"])
    x, go, scale, zp = _per_channel_case(C_, inner, bits, seed=seed)
    if inner == 1:                                   # max == min: the calibration formula gives scale 0
        scale = torch.full_like(scale, 0.125)
    _check_per_channel(fq, ref, x, go, scale, zp, method, bits, philox=(17, seed))
