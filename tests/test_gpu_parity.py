"""GPU parity: the sm_100a kernels (through the C ABI) vs the oracle.

Bars (BASELINE.json north_star): fake-quant outputs and integer codes bit-exact;
scale gradients within 1e-5 relative.  Input gradients are bit-exact too for
STE / LSQ / EWGS; AEWGS input gradients depend on per-channel fp32 means and are
held to 1e-5 relative.

Scale-gradient rule.  d/ds is a sum of N signed terms that the reference
accumulates in fp32 in three separately-rounded pieces which nearly cancel
(SURVEY.md §7 "Hard parts"); its own result moves by more than 1e-5 relative
with the summation order once codes are wide or N is large.  A value passes if
    |ours - ref32| <= 1e-5 * |ref32|                       (the stated bar), or
    |ours - exact| <= max(1e-5 * |exact|, |ref32 - exact|)  (never less accurate
                                                            than the reference)
where `exact` is the fp64 sum of the reference's own fp32 per-element terms.
"""
import math
import zlib

import pytest
import torch

from tests import helpers as H
from oracle import fq_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-5


@pytest.fixture(scope="module")
def fq():
    import mhaq_b200
    return mhaq_b200


# ---------------------------------------------------------------------------
# golden fixtures (outputs of the live reference)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", H.golden_names("act_"))
def test_act_golden(fq, name):
    c = H.load_golden(name)
    r = H.run_act_case(c, fq.fake_quant, device="cuda")
    H.assert_bit_exact(r["y"], c["y"], "y")
    H.assert_bit_exact(r["gx"], c["gx"], "gx")
    ex = H.exact_act_param_grads(c)
    for k in ("g_log_act_s", "g_log_act_q", "g_act_b"):
        if k in c:
            H.assert_param_grad(r[k], c[k], ex[k], REL, k, abs_floor=2e-6)
    # two-call API (codes) and eval-mode extras in one pass
    dev = "cuda"
    # parameter prep on the CPU (as the fixture's generator did), then moved: CUDA exp2f
    # and the CPU's exp2 may differ in the last bit
    s = torch.exp2(torch.tensor([float(c["log_act_s"])]))
    q = torch.exp2(torch.tensor([float(c["log_act_q"])]))
    b = torch.tensor([float(c["act_b"])])
    hi = (b + q - s).to(dev)
    s, b = s.to(dev), b.to(dev)
    y, codes, mm = fq.quantize_eval(c["x"].to(dev), s, b, b, hi, want_y=True, want_codes=True)
    H.assert_bit_exact(codes, c["codes"], "codes")
    H.assert_bit_exact(y, c["y"], "y(eval)")
    mm = mm.cpu()
    assert mm[2].item() == 0
    bw = torch.log2(mm[1] - mm[0] + 1)
    H.assert_bit_exact(bw, c["bw"], "bw")


@pytest.mark.parametrize("name", H.golden_names("w_"))
def test_weight_golden(fq, name):
    c = H.load_golden(name)
    r = H.run_weight_case(c, fq.fake_quant, device="cuda")
    H.assert_bit_exact(r["wq"], c["wq"], "wq")
    if c["method"] == "AEWGS":
        H.assert_close_rel(r["g_weight"], c["g_weight"], REL, "g_weight", abs_floor=1e-6)
    else:
        # identical except for the amin-scattered zero-point term, which is pure
        # fp32 rounding noise (sum(go) - sum(g_u)) in the reference
        H.assert_close_rel(r["g_weight"], c["g_weight"], REL, "g_weight", abs_floor=2e-5)
    H.assert_close_rel(r["g_log_wght_s"], c["g_log_wght_s"], REL, "g_log_wght_s", abs_floor=2e-6)
    if "bq" in c:
        H.assert_bit_exact(r["bq"], c["bq"], "bq")
        H.assert_close_rel(r["g_bias"], c["g_bias"], REL, "g_bias", abs_floor=2e-5)


def test_quantizer_codes_golden(fq):
    c = H.load_golden("quantizer_codes_lsq", device="cuda")
    x = c["x"].clone().requires_grad_(True)
    scale = c["scale"].clone().requires_grad_(True)
    zp = c["zp"].clone().requires_grad_(True)
    codes = fq.quantize_codes(x, scale, zp, -math.inf, math.inf, method="LSQ")
    codes.backward(c["gcodes"])
    H.assert_bit_exact(codes, c["codes"], "codes")
    H.assert_bit_exact(x.grad, c["gx"], "gx")
    H.assert_close_rel(scale.grad, c["g_scale"], REL, "g_scale", abs_floor=2e-6)
    H.assert_close_rel(zp.grad, c["g_zp"], REL, "g_zp", abs_floor=2e-6)


# ---------------------------------------------------------------------------
# seeded random cases vs the oracle, raw-parameter level
# ---------------------------------------------------------------------------
def _leafs(shape_x, per_channel, bits, clip, seed, device):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape_x, generator=g)
    go = torch.randn(shape_x, generator=g)
    if per_channel:
        C = shape_x[0]
        pshape = (C,) + (1,) * (len(shape_x) - 1)
        flat = x.reshape(C, -1)
        mn, mx = flat.amin(1).reshape(pshape), flat.amax(1).reshape(pshape)
    else:
        pshape = (1,)
        mn, mx = x.min().reshape(1), x.max().reshape(1)
    if clip:
        lo = torch.full(pshape, -2.0) + 0.1 * torch.rand(pshape, generator=g)
        q = 4.0
        scale = torch.full(pshape, q / 2 ** bits) * (1 + 0.05 * torch.rand(pshape, generator=g))
        hi = lo + q - scale
        zp = lo.clone()
    else:
        lo = hi = None
        scale = (mx - mn) / (2 ** bits - 1)
        zp = mn.clone()
    r = torch.randint(0, 2, shape_x, generator=g).float() - 0.5
    mk = lambda t: None if t is None else t.to(device).clone().requires_grad_(True)
    return x, go, r, scale, zp, lo, hi, mk


def _run(fn, x, go, r, scale, zp, lo, hi, mk, method, device, expand=False):
    xs = x.to(device).clone().requires_grad_(True)
    P = [scale, zp, lo, hi]
    if expand:  # per-element parameter gradients: fp32 terms, summed in fp64 below
        P = [None if p is None else p.expand(x.shape).contiguous() for p in P]
    s_, z_, l_, h_ = [mk(p) for p in P]
    y = fn(xs, s_, z_, -math.inf if l_ is None else l_, math.inf if h_ is None else h_,
           method=method, noise=None if method == "LSQ" else r.to(device))
    y.backward(go.to(device))
    grads = [None if p is None else p.grad for p in (s_, z_, l_, h_)]
    return y.detach(), xs.grad, grads


def _exact(grads_full, like):
    out = []
    for gfull, p in zip(grads_full, like):
        if gfull is None:
            out.append(None)
            continue
        g64 = gfull.double().cpu()
        if p.numel() == 1:
            out.append(g64.sum().reshape(p.shape))
        else:
            out.append(g64.reshape(p.shape[0], -1).sum(1).reshape(p.shape))
    return out


def _check_param_grad(ours, ref32, exact, what, abs_floor):
    H.assert_param_grad(ours, ref32, exact, REL, what, abs_floor)


CASES = [
    # shape, per_channel, bits, clip, method
    ((2, 3, 20, 17), False, 4, True, "STE"),          # ragged per-tensor, one task
    ((4, 16, 32, 32), False, 4, True, "STE"),         # 65536 = 16 sub-tiles
    ((8, 16, 56, 56), False, 2, True, "STE"),         # 401408: several tasks + ragged tail
    ((3, 5, 7, 11), False, 8, True, "STE"),           # odd sizes -> scalar path
    ((64, 64, 3, 3), True, 4, False, "STE"),          # weight rows of 576
    ((64, 64, 3, 3), True, 3, False, "LSQ"),
    ((64, 64, 3, 3), True, 2, False, "AEWGS"),
    ((64, 64, 3, 3), True, 2, False, "EWGS"),
    ((32, 50, 3, 3), True, 2, False, "LSQ"),          # rows of 450 (RFDN): not a multiple of 4
    ((32, 50, 3, 3), True, 1, False, "AEWGS"),
    ((16, 512, 3, 3), True, 4, False, "STE"),         # rows of 4608: two sub-tiles per row
    ((8, 40000), True, 5, True, "STE"),               # long rows: several tasks per row, clipped
    ((1, 70000), False, 6, False, "LSQ"),             # per-tensor, no clamp
    ((512,), False, 4, False, "STE"),
    ((1, 50, 64, 256), False, 2, True, "STE"),        # RFDN-style activations (config 5), W2A2
    ((4, 12, 41, 41), False, 2, True, "STE"),         # RFDN's 41x41 ESA maps: 80 688 elements, ragged
    ((12, 12, 3, 3), True, 2, False, "LSQ"),          # RFDN's smallest weight: rows of 108
    ((25, 50, 3, 3), True, 2, False, "LSQ"),
]


@pytest.mark.parametrize("shape,per_channel,bits,clip,method", CASES)
def test_random_vs_oracle(fq, shape, per_channel, bits, clip, method):
    seed = zlib.crc32(repr((shape, per_channel, bits, clip, method)).encode()) % 10000
    x, go, r, scale, zp, lo, hi, mk = _leafs(shape, per_channel, bits, clip, seed, "cpu")
    mk_cpu = lambda t: None if t is None else t.clone().requires_grad_(True)
    mk_gpu = lambda t: None if t is None else t.to("cuda").clone().requires_grad_(True)
    y_o, gx_o, g_o = _run(O.fake_quant, x, go, r, scale, zp, lo, hi, mk_cpu, method, "cpu")
    _, _, g_full = _run(O.fake_quant, x, go, r, scale, zp, lo, hi, mk_cpu, method, "cpu", expand=True)
    g_e = _exact(g_full, [scale, zp, lo, hi])
    if method == "AEWGS":   # expanding the scale changes reduce_to_shape's dims: no exact form
        g_e = [None if t is None else t.detach().double().cpu() for t in g_o]
    y_g, gx_g, g_g = _run(fq.fake_quant, x, go, r, scale, zp, lo, hi, mk_gpu, method, "cuda")
    H.assert_bit_exact(y_g, y_o, "y")
    if method == "AEWGS":
        H.assert_close_rel(gx_g, gx_o, REL, "gx", abs_floor=1e-7)
    else:
        H.assert_bit_exact(gx_g, gx_o, "gx")
    n = x.numel() if not per_channel else x.numel() // shape[0]
    floor = 2e-7 * math.sqrt(n) * 4      # fp32 rounding noise of an N-term sum of O(1) terms
    for nm, a, b, e in zip(("g_scale", "g_zp", "g_lo", "g_hi"), g_g, g_o, g_e):
        if b is None:
            continue
        _check_param_grad(a, b, e, nm, floor if nm != "g_scale" else floor * 0.25)


@pytest.mark.parametrize("bits", [1, 2, 3, 4, 5, 6, 7, 8])
def test_bits_sweep_forward_bit_exact(fq, bits):
    """BASELINE config 2 parameterisation: act_b=-2, log_act_q=2, log_act_s=2-bits."""
    torch.manual_seed(bits)
    x = torch.randn(1 << 20)
    s = torch.exp2(torch.tensor([2.0 - bits]))
    b = torch.tensor([-2.0])
    hi = b + 4.0 - s
    y_o = O.fake_quant(x, s, b, b, hi)
    c_o = O.quantize(x, s, b, b, hi)
    y_g, c_g, mm = fq.quantize_eval(x.cuda(), s.cuda(), b.cuda(), b.cuda(), hi.cuda(),
                                    want_y=True, want_codes=True)
    H.assert_bit_exact(y_g, y_o, "y")
    H.assert_bit_exact(c_g, c_o, "codes")
    assert mm[0].item() == c_o.min().item() and mm[1].item() == c_o.max().item()
    assert c_o.max().item() <= 2 ** bits - 1


def test_special_values(fq):
    """NaN / inf / signed zero / exact ties / values on the clip bounds."""
    s = torch.tensor([0.25])
    b = torch.tensor([-1.0])
    hi = torch.tensor([2.0])
    vals = [float("nan"), float("inf"), -float("inf"), 0.0, -0.0, -1.0, 2.0, 2.0000002, -1.0000001,
            -1.0 + 0.125, -1.0 + 0.375, -1.0 + 0.625, 1e-30, -1e30, 1.9999999]
    x = torch.tensor(vals * 40)[:512].clone()
    go = torch.ones_like(x)
    for lo_, hi_ in ((b, hi), (None, None)):
        xo = x.clone().requires_grad_(True)
        y_o = O.fake_quant(xo, s, b, -math.inf if lo_ is None else lo_, math.inf if hi_ is None else hi_,
                           method="LSQ")
        y_o.backward(go)
        xg = x.cuda().requires_grad_(True)
        y_g = fq.fake_quant(xg, s.cuda(), b.cuda(), -math.inf if lo_ is None else lo_.cuda(),
                            math.inf if hi_ is None else hi_.cuda(), method="LSQ")
        y_g.backward(go.cuda())
        H.assert_bit_exact(y_g, y_o, "y")
        H.assert_bit_exact(xg.grad, xo.grad, "gx")


def test_empty_and_tiny(fq):
    s = torch.tensor([0.5], device="cuda", requires_grad=True)
    z = torch.tensor([0.0], device="cuda")
    x = torch.empty(0, 4, device="cuda", requires_grad=True)
    y = fq.fake_quant(x, s, z)
    assert y.shape == (0, 4)
    y.sum().backward()
    assert s.grad is not None and s.grad.item() == 0.0
    x1 = torch.tensor([0.74], device="cuda", requires_grad=True)
    y1 = fq.fake_quant(x1, s, z, method="LSQ")
    assert y1.item() == 0.5
    y1.backward(torch.ones_like(y1))
    assert x1.grad.item() == 1.0


def test_cpu_tensor_is_an_error(fq):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fq.fake_quant(torch.randn(8), torch.tensor([0.5]), torch.tensor([0.0]))


# ---------------------------------------------------------------------------
# in-kernel Philox noise
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("rows,inner", [(1, 100000), (3, 16384 + 520), (5, 450), (2, 4096)])
def test_philox_stream_matches_numpy(fq, rows, inner):
    like = torch.empty(rows, inner, device="cuda")
    sc = torch.empty(rows, 1, device="cuda") if rows > 1 else None
    r = fq.philox_noise(like, sc, seed=0x1234567890ABCDEF, offset=77).cpu().numpy()
    ref = O.philox_noise(rows, inner, 0x1234567890ABCDEF, 77)
    assert (r == ref).all()
    assert abs(float(r.mean())) < 0.02


@pytest.mark.parametrize("shape,per_channel,clip", [((8, 16, 56, 56), False, True),
                                                     ((64, 64, 3, 3), True, False)])
def test_fused_noise_equals_explicit_noise(fq, shape, per_channel, clip):
    """Drawing r in-kernel must give the SAME gradients as feeding the materialised r."""
    x, go, _, scale, zp, lo, hi, _ = _leafs(shape, per_channel, 4, clip, 5, "cuda")
    mk = lambda t: None if t is None else t.cuda().clone().requires_grad_(True)
    seed, off = 99, 1234
    r = fq.philox_noise(x.cuda(), scale.cuda() if per_channel else None, seed=seed, offset=off)

    def run(noise, philox):
        xs = x.cuda().requires_grad_(True)
        s_, z_, l_, h_ = mk(scale), mk(zp), mk(lo), mk(hi)
        y = fq.fake_quant(xs, s_, z_, -math.inf if l_ is None else l_, math.inf if h_ is None else h_,
                          method="STE", noise=noise, philox=philox)
        y.backward(go.cuda())
        return xs.grad, s_.grad
    gx_a, gs_a = run(r, None)
    gx_b, gs_b = run(None, (seed, off))
    assert torch.equal(gx_a, gx_b)
    assert torch.equal(gs_a, gs_b), (gs_a - gs_b).abs().max()


def test_deterministic(fq):
    x, go, r, scale, zp, lo, hi, _ = _leafs((8, 16, 56, 56), False, 4, True, 11, "cuda")
    mk = lambda t: t.cuda().clone().requires_grad_(True)
    outs = []
    for _ in range(3):
        xs = x.cuda().requires_grad_(True)
        s_, z_, l_, h_ = mk(scale), mk(zp), mk(lo), mk(hi)
        y = fq.fake_quant(xs, s_, z_, l_, h_, method="STE", philox=(1, 2))
        y.backward(go.cuda())
        outs.append((y.detach(), xs.grad, s_.grad, z_.grad, l_.grad, h_.grad))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)


def test_flat_backward_record_handover_is_stateless(fq):
    """The single-launch per-tensor backward hands its per-block records to block 0 through
    self-validating words in the stream's ticket buffer (include/mhaq_fq.h).  Back-to-back
    launches of different sizes (different grids, so different record counts) on one stream, then
    replays of one captured launch: every result equals the first run of that size bit for bit,
    and the buffer is zero again after every launch."""
    from mhaq_b200 import ops
    sizes = [2048 * 9, 6422528, 2048 * 592 * 3 + 2048 * 7 + 4, 4 << 20, 2044, 1 << 24, 12845056]
    b = torch.tensor([-2.0], device="cuda")
    s = torch.tensor([0.25], device="cuda")
    hi = b + 4.0 - s
    data = {}
    for n in sizes:
        g = torch.Generator(device="cuda").manual_seed(n % 1009)
        data[n] = (torch.randn(n, device="cuda", generator=g) * 1.5, torch.randn(n, device="cuda", generator=g))

    def run(n):
        x, go = data[n]
        assert ops.lib.mhaq_fq_bwd_single_launch(1, n, 1, 0, 0) == 1
        xs = x.clone().requires_grad_(True)
        s_, b_, h_ = (t.clone().requires_grad_(True) for t in (s, b, hi))
        fq.fake_quant(xs, s_, b_, b_, h_, method="STE", philox=(3, 4)).backward(go)
        return xs.grad, s_.grad, b_.grad, h_.grad

    first = {n: run(n) for n in sizes}
    for rep in range(3):
        for n in sizes[::-1] if rep % 2 else sizes:
            for a, c in zip(first[n], run(n)):
                assert torch.equal(a, c), (n, rep)
    torch.cuda.synchronize()
    tk = ops._arena(data[sizes[0]][0]).tickets
    assert tk is not None and bool((tk == 0).all()), "the ticket / record buffer must be left zero"

    # one captured launch replayed many times (the training step's situation)
    n = 6422528
    x, go = data[n]
    L = ops._Launch(x, s, b, b, hi)
    gx = torch.empty_like(x)
    out = torch.zeros(4, 1, device="cuda")
    ws = ops._workspace(x, L.geo)
    tk2 = torch.zeros(ops.lib.mhaq_fq_ticket_count(1, n, 1), dtype=torch.int32, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())

    def launch():
        ops.check(ops.lib.mhaq_fq_bwd_fused_f32(
            go.data_ptr(), x.data_ptr(), gx.data_ptr(), *L.params(), 1, n, 1, 0, 0, None, 3, 4, None, None,
            ws.data_ptr(), tk2.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
            out[3].data_ptr(), ops._stream()), "bwd_fused")
    with torch.cuda.stream(side):
        launch()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        launch()
    for _ in range(50):
        out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(gx, first[n][0])
        assert torch.equal(out[0], first[n][1]) and torch.equal(out[3], first[n][3])
    assert bool((tk2 == 0).all())


def test_graph_capture_gets_its_own_scratch_arena(fq):
    """A captured backward must not share the ticket / record buffer of the stream it was captured
    from: the graph is replayed on the default stream here while eager calls keep running on a
    side stream, concurrently.  Both keep producing the eager result bit for bit."""
    from mhaq_b200 import ops
    n = 6422528
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(n, device="cuda", generator=g) * 1.5
    go = torch.randn(n, device="cuda", generator=g)
    b = torch.tensor([-2.0], device="cuda")
    s = torch.tensor([0.25], device="cuda")
    hi = b + 4.0 - s

    def run():
        xs = x.clone().requires_grad_(True)
        s_, b_, h_ = (t.clone().requires_grad_(True) for t in (s, b, hi))
        y = fq.fake_quant(xs, s_, b_, b_, h_, method="STE", philox=(3, 4))
        return torch.autograd.grad(y, (xs, s_, b_, h_), go)

    want = run()
    torch.cuda.synchronize()
    sx = x.clone().requires_grad_(True)
    sp = [t.clone().requires_grad_(True) for t in (s, b, hi)]
    before = set(ops._arenas)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = fq.fake_quant(sx, sp[0], sp[1], sp[1], sp[2], method="STE", philox=(3, 4))
        got = torch.autograd.grad(y, (sx, sp[0], sp[1], sp[2]), go)
    new = [k for k in set(ops._arenas) - before if k[2] != 0]
    assert new, "the capture did not get an arena keyed by its capture id"
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    for _ in range(25):
        graph.replay()
        with torch.cuda.stream(side):
            eager = run()
        torch.cuda.synchronize()
        for a, c, e in zip(want, got, eager):
            assert torch.equal(a, c), "graph replay"
            assert torch.equal(a, e), "eager call next to the replay"


# ---------------------------------------------------------------------------
# full-size checks (BASELINE config 2): the oracle's ATen chain runs on the GPU
# itself as the checker; plus size-independent properties
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("log2n,bits", [(26, 4), (28, 8), (30, 4)])     # the single-launch flat kernel at every size
def test_full_size_vs_oracle_on_device(fq, log2n, bits):
    n = 1 << log2n
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, device="cuda", generator=g)
    g1 = torch.Generator(device="cuda").manual_seed(1)
    go = torch.randn(n, device="cuda", generator=g1)
    b = torch.tensor([-2.0], device="cuda")
    s = torch.exp2(torch.tensor([2.0 - bits], device="cuda"))
    hi = b + 4.0 - s
    r = fq.philox_noise(x, None, seed=3, offset=4)

    def run(fn, **kw):
        xs = x.clone().requires_grad_(True)
        s_, b_, h_ = s.clone().requires_grad_(True), b.clone().requires_grad_(True), hi.clone().requires_grad_(True)
        y = fn(xs, s_, b_, b_, h_, method="STE", **kw)
        y.backward(go)
        return y.detach(), xs.grad, s_.grad, b_.grad, h_.grad
    y_o, gx_o, gs_o, gb_o, gh_o = run(O.fake_quant, noise=r)
    y_g, gx_g, gs_g, gb_g, gh_g = run(fq.fake_quant, philox=(3, 4))
    assert torch.equal(y_g, y_o)
    assert torch.equal(gx_g, gx_o)
    del y_o, gx_o
    # exact (fp64) value of the scale gradient from fp32 per-element terms
    v = (torch.clamp(x, b, hi) - b) / s
    e = torch.round(v) - v
    inr = (x >= b) & (x <= hi)
    exact_s = (go.double() * e.double()).sum() + ((3.0 ** -0.5) * (go * s)).double().mul(r.double()).sum()
    exact_h = (go * s / s).double()[x > hi].sum()
    _check_param_grad(gs_g, gs_o, exact_s.cpu().reshape(1), "g_scale", 1e-4)
    _check_param_grad(gh_g, gh_o, exact_h.cpu().reshape(1), "g_hi", 1e-4)
    # properties: codes in range and integral, clipped elements pass no gradient
    _, codes, mm = fq.quantize_eval(x, s, b, b, hi, want_y=False, want_codes=True)
    assert mm[0].item() >= 0 and mm[1].item() <= 2 ** bits - 1 and mm[2].item() == 0
    assert torch.equal(codes, codes.round())
    assert (gx_g[~inr] == 0).all()
    # linearity of the backward in go (exact for a power-of-two factor)
    xs = x.clone().requires_grad_(True)
    y2 = fq.fake_quant(xs, s, b, b, hi, method="STE", philox=(3, 4))
    y2.backward(go * 4.0)
    assert torch.equal(xs.grad, gx_g * 4.0)


# ---------------------------------------------------------------------------
# row statistics (amin / amax with tie counts) and their backward
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("rows,inner", [(64, 576), (32, 450), (7, 4608), (3, 10), (1, 100000)])
def test_rowstat_matches_torch_amin_amax_and_backward(fq, rows, inner):
    from mhaq_b200 import ops
    torch.manual_seed(rows * 1000 + inner)
    w = torch.randn(rows, inner)
    w[0, 3] = w[0].min()          # ties for the minimum
    w[0, 5] = w[0].min()
    w[rows - 1, 1] = w[rows - 1].max()    # tie for the maximum
    wg = w.cuda()
    mn, mx, cmn, cmx = ops.row_stats(wg)
    assert torch.equal(mn.cpu(), w.amin(1)) and torch.equal(mx.cpu(), w.amax(1))
    assert torch.equal(cmn.cpu(), (w == w.amin(1, keepdim=True)).sum(1).float())
    assert torch.equal(cmx.cpu(), (w == w.amax(1, keepdim=True)).sum(1).float())
    # backward: torch's amin/amax split the gradient evenly among ties
    gmn, gmx = torch.randn(rows), torch.randn(rows)
    gx = torch.randn(rows, inner)
    wr = w.clone().requires_grad_(True)
    ((wr.amin(1) * gmn).sum() + (wr.amax(1) * gmx).sum() + (wr * gx).sum()).backward()
    out = ops.row_stats_backward(gx.cuda(), wg, mn, cmn, gmn.cuda(), mx, cmx, gmx.cuda())
    H.assert_close_rel(out, wr.grad, 1e-6, "rowstat backward", abs_floor=1e-7)
    out2 = ops.row_stats_backward(gx.cuda(), wg, mn, cmn, gmn.cuda())
    wr.grad = None
    ((wr.amin(1) * gmn).sum() + (wr * gx).sum()).backward()
    H.assert_close_rel(out2, wr.grad, 1e-6, "rowstat backward (min only)", abs_floor=1e-7)


@pytest.mark.parametrize("rows,inner", [(1, 4096 * 3 + 4), (1, 4099), (5, 450), (3, 16384 + 512), (7, 1), (2, 8192)])
def test_no_out_of_bounds_writes(fq, rows, inner):
    """compute-sanitizer is closed on this pool, so bounds are checked directly: outputs are
    carved out of sentinel-filled buffers (16-byte aligned AND deliberately misaligned) and the
    sentinels must survive forward and backward for ragged and exact-multiple sizes."""
    from mhaq_b200 import ops
    n = rows * inner
    pad = 64
    for shift in (0, 1):                       # 1 -> data_ptr not 16-byte aligned -> scalar path
        torch.manual_seed(n + shift)
        def carve():
            buf = torch.full((n + 2 * pad + shift,), 1234.5, device="cuda")
            return buf, buf[pad + shift: pad + shift + n].view(rows, inner)
        xb, x = carve(); x.copy_(torch.randn(rows, inner))
        gb, go = carve(); go.copy_(torch.randn(rows, inner))
        yb, y = carve(); cb, codes = carve(); gxb, gx = carve()
        s = torch.full((rows, 1), 0.3, device="cuda") if rows > 1 else torch.tensor([0.3], device="cuda")
        zp = torch.full_like(s, -1.0)
        L = ops._Launch(x, s, zp, zp, zp + 3.0)
        geo = L.geo
        st = ops._stream()
        ops.check(ops.lib.mhaq_fq_fwd_f32(x.data_ptr(), y.data_ptr(), codes.data_ptr(), *L.params(),
                                          geo.n_rows, geo.n_inner, geo.n_ch, None, st), "fwd")
        ws = ops._workspace(x, geo); tk = ops._tickets(x, geo)
        out = torch.empty(4, geo.n_ch, device="cuda")
        ops.check(ops.lib.mhaq_fq_bwd_f32(go.data_ptr(), x.data_ptr(), gx.data_ptr(), *L.params(),
                                          geo.n_rows, geo.n_inner, geo.n_ch, 0, 0, None, 1, 2, None, None,
                                          ws.data_ptr(), st), "bwd")
        ops.check(ops.lib.mhaq_fq_bwd_finalize_f32(ws.data_ptr(), tk.data_ptr(), *L.params(), geo.n_rows,
                                                   geo.n_inner, geo.n_ch, out[0].data_ptr(), out[1].data_ptr(),
                                                   out[2].data_ptr(), out[3].data_ptr(), st), "fin")
        torch.cuda.synchronize()
        for name, buf in (("y", yb), ("codes", cb), ("gx", gxb), ("x", xb), ("go", gb)):
            assert bool((buf[: pad + shift] == 1234.5).all()) and bool((buf[pad + shift + n:] == 1234.5).all()), \
                f"{name}: sentinel overwritten (rows={rows}, inner={inner}, shift={shift})"
        assert bool((tk == 0).all()), "tickets must be restored to zero"
        # and the carved (possibly misaligned) run agrees with the oracle
        yo = O.fake_quant(x.cpu(), s.cpu(), zp.cpu(), zp.cpu(), zp.cpu() + 3.0)
        H.assert_bit_exact(y, yo, "y")


def test_per_tensor_aewgs_dim0_statistics_quirk(fq):
    """Reference quirk 9: scale of shape (1,) -> AEWGS statistics are averaged over dim 0 only
    (gdnsq.py:150-152), i.e. delta is per inner position."""
    torch.manual_seed(21)
    w = torch.randn(12, 5, 3, 3) * 0.4
    go = torch.randn_like(w)
    r = torch.randint(0, 2, w.shape).float() - 0.5
    log_s = torch.tensor([-2.0])

    def run(fn, dev):
        wr = w.to(dev).clone().requires_grad_(True)
        ls = log_s.to(dev).clone().requires_grad_(True)
        s = torch.exp2(ls)
        y = fn(wr, s, wr.amin(), -math.inf, math.inf, method="AEWGS", noise=r.to(dev))
        y.backward(go.to(dev))
        return y.detach(), wr.grad, ls.grad
    y_o, gw_o, gs_o = run(O.fake_quant, "cpu")
    y_g, gw_g, gs_g = run(fq.fake_quant, "cuda")
    H.assert_bit_exact(y_g, y_o, "y")
    H.assert_close_rel(gw_g, gw_o, REL, "g_weight", abs_floor=2e-6)
    H.assert_close_rel(gs_g, gs_o, REL, "g_log_s", abs_floor=2e-5)


@pytest.mark.parametrize("shape,per_channel,clip", [((24, 20000), True, False), ((1, 1 << 20), False, True),
                                                     ((16, 8, 3, 3), True, False)])
def test_host_streaming_api_matches_device_api(fq, shape, per_channel, clip):
    """The pipelined host-buffer entry (bench `e2e`) must give the device API's results."""
    from mhaq_b200.host import fake_quant_fwd_bwd_host
    x, go, _, scale, zp, lo, hi, _ = _leafs(shape, per_channel, 4, clip, 17, "cpu")
    dev = "cuda"
    P = [None if p is None else p.to(dev) for p in (scale, zp, lo, hi)]
    # device API (fused Philox with the same per-chunk streams is not comparable: use LSQ)
    xs = x.to(dev).requires_grad_(True)
    Pd = [None if p is None else p.clone().requires_grad_(True) for p in P]
    y = fq.fake_quant(xs, Pd[0], Pd[1], -math.inf if Pd[2] is None else Pd[2],
                      math.inf if Pd[3] is None else Pd[3], method="LSQ")
    y.backward(go.to(dev))
    hy, hgx, grads = fake_quant_fwd_bwd_host(x.pin_memory(), go.pin_memory(), P[0], P[1], P[2], P[3],
                                             method="LSQ", chunks=5)
    torch.cuda.synchronize()
    H.assert_bit_exact(hy, y, "y")
    H.assert_bit_exact(hgx, xs.grad, "gx")
    for k, p in zip(("scale", "zero_point", "min_val", "max_val"), Pd):
        if p is not None:
            H.assert_close_rel(grads[k], p.grad, 1e-5, k, abs_floor=1e-5)


# ---------------------------------------------------------------------------
# channels_last (NHWC) tensors are walked in storage order, without layout copies
# ---------------------------------------------------------------------------
def _cl(t):
    return t.contiguous(memory_format=torch.channels_last if t.dim() == 4 else torch.channels_last_3d)


@pytest.mark.parametrize("shape", [(4, 16, 9, 7), (2, 3, 5, 5), (8, 16, 56, 56), (2, 4, 3, 6, 5)])
@pytest.mark.parametrize("go_layout", ["same", "row_major"])
def test_channels_last_activation_equals_row_major(fq, shape, go_layout):
    """Per-tensor parameters: a channels_last activation is quantized in place (output and input
    gradient keep the NHWC strides) and every element gets exactly the row-major result."""
    x, go, r, scale, zp, lo, hi, mk = _leafs(shape, False, 4, True, 23, "cuda")
    y0, gx0, g0 = _run(fq.fake_quant, x, go, r, scale, zp, lo, hi, mk, "STE", "cuda")
    xs = _cl(x.cuda()).requires_grad_(True)
    P = [mk(p) for p in (scale, zp, lo, hi)]
    y = fq.fake_quant(xs, *P, method="STE", noise=_cl(r.cuda()))
    assert y.stride() == xs.stride()
    y.backward(_cl(go.cuda()) if go_layout == "same" else go.cuda())
    assert xs.grad.stride() == xs.stride()
    H.assert_bit_exact(y, y0, "y")
    H.assert_bit_exact(xs.grad, gx0, "gx")
    for p, g, k in zip(P, g0, ("scale", "zp", "lo", "hi")):
        H.assert_close_rel(p.grad, g, REL, k, abs_floor=2e-5)
    # fused NoisyAct entry (log-domain parameters) on the same layout
    ls = torch.tensor([-2.0], device="cuda", requires_grad=True)
    lq = torch.tensor([2.0], device="cuda", requires_grad=True)
    ab = torch.tensor([-2.0], device="cuda", requires_grad=True)
    outs = []
    for xin, gin in ((x.cuda(), go.cuda()), (_cl(x.cuda()), _cl(go.cuda()))):
        for p in (ls, lq, ab):
            p.grad = None
        xin = xin.requires_grad_(True)
        ya = fq.ops.act_fake_quant(xin, ls, lq, ab, method="LSQ")
        ya.backward(gin)
        assert ya.stride() == xin.stride() and xin.grad.stride() == xin.stride()
        outs.append((ya.detach(), xin.grad, ls.grad.clone(), lq.grad.clone(), ab.grad.clone()))
    H.assert_bit_exact(outs[1][0], outs[0][0], "act y")
    H.assert_bit_exact(outs[1][1], outs[0][1], "act gx")
    for i, k in ((2, "log_act_s"), (3, "log_act_q"), (4, "act_b")):
        H.assert_close_rel(outs[1][i], outs[0][i], REL, k, abs_floor=2e-5)


@pytest.mark.parametrize("shape,method", [((64, 64, 3, 3), "STE"), ((32, 50, 3, 3), "LSQ"),
                                          ((16, 512, 3, 3), "AEWGS")])
def test_channels_last_weight_equals_row_major(fq, shape, method):
    """Per-channel (dim 0) weights in channels_last: each output channel is still one dense
    block; row statistics, fake-quant, gradients and the amin scatter act on it in place."""
    g = torch.Generator().manual_seed(5)
    w = torch.randn(shape, generator=g).cuda()
    go = torch.randn(shape, generator=g).cuda()
    r = (torch.randint(0, 2, shape, generator=g).float() - 0.5).cuda()
    log_s = (torch.full((shape[0], 1, 1, 1), -3.0) + 0.2 * torch.rand((shape[0], 1, 1, 1), generator=g)).cuda()
    res = []
    for conv in (lambda t: t, _cl):
        wl = conv(w).clone(memory_format=torch.preserve_format).requires_grad_(True)
        ls = log_s.clone().requires_grad_(True)
        noise = None if method == "LSQ" else conv(r)
        wq, mn, mx = fq.ops.weight_fake_quant_log(wl, ls, method=method, noise=noise)
        assert wq.stride() == wl.stride()
        loss = (wq * conv(go)).sum() + (mx - mn).sum()
        loss.backward()
        assert wl.grad.stride() == wl.stride()
        res.append((wq.detach(), wl.grad, ls.grad, mn.detach(), mx.detach()))
    H.assert_bit_exact(res[1][0], res[0][0], "wq")
    H.assert_bit_exact(res[1][3], res[0][3], "row_min")
    H.assert_bit_exact(res[1][4], res[0][4], "row_max")
    if method == "AEWGS":
        H.assert_close_rel(res[1][1], res[0][1], REL, "g_weight", abs_floor=2e-6)
    else:
        H.assert_close_rel(res[1][1], res[0][1], REL, "g_weight", abs_floor=2e-5)
    H.assert_close_rel(res[1][2], res[0][2], REL, "g_log_wght_s", abs_floor=2e-5)


# ---------------------------------------------------------------------------
# row-resident fused weight kernels (mhaq_fq_wrow_*): one launch each way
# ---------------------------------------------------------------------------
WROW_SHAPES = [(64, 64, 3, 3), (32, 50, 3, 3), (16, 512, 3, 3), (12, 12, 3, 3), (24, 40), (7, 1, 3, 3),
               (3, 16384), (5, 4099)]


@pytest.mark.parametrize("shape", WROW_SHAPES)
@pytest.mark.parametrize("method", ["STE", "LSQ", "EWGS"])
def test_fused_weight_rows_match_oracle_and_streaming_path(fq, shape, method):
    """(wq, row_min, row_max, log2(max-min+2^log_s)) and every gradient of the fused per-row
    kernels against (a) the oracle's op-for-op restatement of NoisyConv2d.forward's weight lines
    plus ModelHelper's expression under torch autograd on the CPU and (b) the streaming kernels."""
    ops = fq.ops
    g = torch.Generator().manual_seed(zlib.crc32(repr((shape, method)).encode()))
    w = torch.randn(shape, generator=g)
    w[0].view(-1)[:3] = w[0].min()                  # ties at the row minimum ...
    w[-1].view(-1)[-2:] = w[-1].max()               # ... and at the row maximum
    go = torch.randn(shape, generator=g)
    gl = torch.randn(shape[0], generator=g)
    r = torch.randint(0, 2, shape, generator=g).float() - 0.5
    pshape = (shape[0],) + (1,) * (len(shape) - 1)
    log_s = torch.full(pshape, -3.0) + 0.25 * torch.rand(pshape, generator=g)
    noise = None if method == "LSQ" else r
    dims = tuple(range(1, len(shape)))

    def reference(dev):
        wr = w.to(dev).clone().requires_grad_(True)
        ls = log_s.to(dev).clone().requires_grad_(True)
        wq = O.weight_fake_quant(wr, ls, True, method, noise=None if noise is None else noise.to(dev))
        lr = torch.log2(wr.amax(dims) - wr.amin(dims) + torch.exp2(ls.ravel()))
        ((wq * go.to(dev)).sum() + (lr * gl.to(dev)).sum()).backward()
        return wq.detach(), lr.detach(), wr.grad, ls.grad

    # exp2 is evaluated inside the kernels (CUDA exp2f == torch's CUDA exp2 kernel, while the
    # CPU's vectorised exp2 may differ in the last bit): the op-for-op oracle therefore runs on
    # the device for fractional log-scales ...
    wq_o, lr_o, gw_o, gs_o = reference("cuda")
    assert ops.weight_rows_fusable(w.cuda(), log_s.cuda(), method)
    # The weight gradient is the per-element quantizer gradient (deterministic: bit-exact) plus,
    # on the row's min / max elements only, the amin / amax backward of the zero-point gradient.
    # That term is sum(go) - sum(g_u) accumulated in fp32 by the reference — rounding noise that
    # grows with the row length — while the kernels sum the per-element differences in fp64.
    n_inner = w[0].numel()
    flat = w.reshape(shape[0], -1)
    ties = ((flat == flat.amin(1, keepdim=True)) | (flat == flat.amax(1, keepdim=True))).reshape(shape)
    tie_floor = max(2e-5, 1.5e-6 * math.sqrt(n_inner))

    # d/d log_wght_s: the reference's fp32 value is itself only good to ~1e-4 on long rows (two
    # big fp32 sums that nearly cancel, see the module docstring), so it is held to the
    # parameter-gradient rule: within 1e-5 of the reference's fp32 value, or at least as close as
    # the reference to `exact` = the fp64 sum of the reference's own fp32 per-element terms
    # (chain rule to the log domain and the range term's own contribution added in fp64).
    def exact_g_log_s(dev, ls_t):
        wd, s64 = w.to(dev), torch.exp2(ls_t.to(dev)).double().cpu().ravel()
        s32 = torch.exp2(ls_t.to(dev))
        zp = wd.amin(dims, keepdim=True)
        ex = H.exact_param_grads(O.fake_quant, wd, go.to(dev), s32, zp, None, None, method,
                                 None if noise is None else noise.to(dev))[0].double().ravel()
        rng = (wd.amax(dims) - wd.amin(dims)).double().cpu().ravel()
        return ex * s64 * math.log(2.0) + gl.double() * s64 / (rng + s64)

    gs_floor = 4e-7 * math.sqrt(n_inner)

    def check_gw(ours, ref, what):
        ours, ref = ours.detach().cpu(), ref.detach().cpu()
        H.assert_bit_exact(ours[~ties], ref[~ties], what + " (off the row extrema)")
        H.assert_close_rel(ours[ties], ref[ties], REL, what + " (row extrema)", abs_floor=tie_floor)

    wr = w.cuda().requires_grad_(True)
    ls = log_s.cuda().requires_grad_(True)
    wq, mn, mx, lr = ops.weight_fake_quant_rows(wr, ls, method=method, noise=None if noise is None else noise.cuda())
    ((wq * go.cuda()).sum() + (lr * gl.cuda()).sum()).backward()
    H.assert_bit_exact(wq, wq_o, "wq")
    H.assert_bit_exact(mn, w.amin(dims), "row_min")
    H.assert_bit_exact(mx, w.amax(dims), "row_max")
    H.assert_close_rel(lr, lr_o, 1e-6, "log_range", abs_floor=1e-6)
    check_gw(wr.grad, gw_o, "g_weight")
    H.assert_param_grad(ls.grad.ravel(), gs_o.ravel(), exact_g_log_s("cuda", log_s), REL, "g_log_wght_s", gs_floor)
    # ... and on the CPU for integer log-scales, where exp2 is exact everywhere
    log_s_frac, log_s = log_s, log_s.round()
    wq_c, lr_c, gw_c, gs_c = reference("cpu")
    wr1 = w.cuda().requires_grad_(True)
    ls1 = log_s.cuda().requires_grad_(True)
    wq1, _, _, lr1 = ops.weight_fake_quant_rows(wr1, ls1, method=method, noise=None if noise is None else noise.cuda())
    ((wq1 * go.cuda()).sum() + (lr1 * gl.cuda()).sum()).backward()
    H.assert_bit_exact(wq1, wq_c, "wq (CPU oracle)")
    H.assert_close_rel(lr1, lr_c, 1e-6, "log_range (CPU oracle)", abs_floor=1e-6)
    check_gw(wr1.grad, gw_c, "g_weight (CPU oracle)")
    H.assert_param_grad(ls1.grad.ravel(), gs_c.ravel(), exact_g_log_s("cpu", log_s), REL, "g_log_wght_s (CPU oracle)", gs_floor)
    log_s = log_s_frac
    # the streaming path + torch autograd for the range term, on the same device
    wr2 = w.cuda().requires_grad_(True)
    ls2 = log_s.cuda().requires_grad_(True)
    wq2, mn2, mx2 = ops.weight_fake_quant_log(wr2, ls2, method=method, noise=None if noise is None else noise.cuda())
    lr2 = torch.log2(mx2 - mn2 + torch.exp2(ls2.ravel()))
    ((wq2 * go.cuda()).sum() + (lr2 * gl.cuda()).sum()).backward()
    H.assert_bit_exact(wq, wq2, "wq vs streaming")
    H.assert_bit_exact(lr, lr2, "log_range vs torch.log2 on the device")
    H.assert_close_rel(wr.grad, wr2.grad, REL, "g_weight vs streaming", abs_floor=2e-6)
    H.assert_close_rel(ls.grad, ls2.grad, REL, "g_log_wght_s vs streaming", abs_floor=2e-6)


@pytest.mark.parametrize("shape", [(32, 50, 3, 3), (16, 512, 3, 3), (3, 16384)])
def test_fused_weight_rows_philox_stream_is_the_streaming_kernels_stream(fq, shape):
    """Same (seed, offset, row, position) -> same noise bit as mhaq_fq_bwd_f32 / mhaq_fq_noise_f32."""
    ops = fq.ops
    g = torch.Generator().manual_seed(11)
    w = torch.randn(shape, generator=g).cuda()
    go = torch.randn(shape, generator=g).cuda()
    log_s = torch.full((shape[0],) + (1,) * (len(shape) - 1), -3.0).cuda()
    r = ops.philox_noise(w, log_s, seed=77, offset=5)
    res = []
    for kw in (dict(philox=(77, 5)), dict(noise=r)):
        ls = log_s.clone().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        wq, _, _, _ = ops.weight_fake_quant_rows(wr, ls, method="STE", **kw)
        wq.backward(go)
        res.append((wr.grad, ls.grad))
    H.assert_bit_exact(res[0][0], res[1][0], "g_weight")
    H.assert_bit_exact(res[0][1], res[1][1], "g_log_wght_s")


def test_fused_weight_rows_channels_last_and_unused_outputs(fq):
    ops = fq.ops
    g = torch.Generator().manual_seed(3)
    w = torch.randn(16, 8, 3, 3, generator=g).cuda()
    log_s = torch.full((16, 1, 1, 1), -3.0).cuda()
    outs = []
    for conv in (lambda t: t, _cl):
        wr = conv(w).clone(memory_format=torch.preserve_format).requires_grad_(True)
        ls = log_s.clone().requires_grad_(True)
        wq, mn, mx, lr = ops.weight_fake_quant_rows(wr, ls, method="LSQ")
        assert wq.stride() == wr.stride()
        lr.sum().backward()                       # only the range term is used (wq's grad is None)
        assert wr.grad.stride() == wr.stride()
        outs.append((wq.detach(), wr.grad, ls.grad))
    H.assert_bit_exact(outs[1][0], outs[0][0], "wq")
    H.assert_bit_exact(outs[1][1], outs[0][1], "g_weight")
    H.assert_bit_exact(outs[1][2], outs[0][2], "g_log_wght_s")
    # against torch autograd of the range term alone
    wr = w.clone().requires_grad_(True)
    ls = log_s.clone().requires_grad_(True)
    torch.log2(wr.amax((1, 2, 3)) - wr.amin((1, 2, 3)) + torch.exp2(ls.ravel())).sum().backward()
    H.assert_close_rel(outs[0][1], wr.grad, REL, "g_weight (range only)", abs_floor=1e-6)
    H.assert_close_rel(outs[0][2], ls.grad, REL, "g_log_wght_s (range only)", abs_floor=1e-6)
    assert not ops.weight_rows_fusable(w, log_s, "AEWGS")
    with pytest.raises(RuntimeError):
        ops.weight_fake_quant_rows(torch.randn(2, 20000, device="cuda"), log_s[:2], method="STE")


# ---------------------------------------------------------------------------
# multi-tensor weight launch (mhaq_fq_wrow_multi_*): identical to the per-layer kernels
# ---------------------------------------------------------------------------
MULTI_SHAPES = [(16, 16, 3, 3), (32, 16, 3, 3), (50, 50, 3, 3), (12, 50, 3, 3), (64, 64, 3, 3), (8, 4608),
                (25, 50, 3, 3), (7, 9)]


@pytest.mark.parametrize("method", ["STE", "LSQ", "EWGS"])
@pytest.mark.parametrize("n_tensors", [8, 53])          # 53 > 32 / 24: several launches per direction
def test_multi_tensor_weight_launch_equals_per_layer_kernels(fq, method, n_tensors):
    ops = fq.ops
    g = torch.Generator().manual_seed(n_tensors)
    shapes = [MULTI_SHAPES[i % len(MULTI_SHAPES)] for i in range(n_tensors)]
    ws = [(torch.randn(s, generator=g) * 0.2).cuda() for s in shapes]
    lss = [(torch.full((s[0],) + (1,) * (len(s) - 1), -4.0) + 0.5 * torch.rand((s[0],) + (1,) * (len(s) - 1), generator=g)).cuda()
           for s in shapes]
    gos = [torch.randn(s, generator=g).cuda() for s in shapes]
    gls = [torch.randn(s[0], generator=g).cuda() for s in shapes]
    for mode in ("explicit", "philox"):
        if method == "LSQ" and mode == "philox":
            continue
        noises = [(torch.randint(0, 2, s, generator=g).float() - 0.5).cuda() for s in shapes] if mode == "explicit" else None
        # per layer
        ref = []
        for i in range(n_tensors):
            w, ls = ws[i].clone().requires_grad_(True), lss[i].clone().requires_grad_(True)
            kw = dict(noise=noises[i]) if noises is not None and method != "LSQ" else dict(philox=(9, 100 + i))
            wq, mn, mx, lr = ops.weight_fake_quant_rows(w, ls, method=method, **kw)
            ((wq * gos[i]).sum() + (lr * gls[i]).sum()).backward()
            ref.append((wq.detach(), mn.detach(), mx.detach(), lr.detach(), w.grad, ls.grad))
        # all at once
        W = [w.clone().requires_grad_(True) for w in ws]
        L = [l.clone().requires_grad_(True) for l in lss]
        res = ops.weight_fake_quant_rows_multi(W, L, method=method, noises=noises, philox=(9, 100))
        loss = sum((wq * go).sum() + (lr * gl).sum() for (wq, mn, mx, lr), go, gl in zip(res, gos, gls))
        loss.backward()
        for i, ((wq, mn, mx, lr), r) in enumerate(zip(res, ref)):
            for a, b, nm in zip((wq, mn, mx, lr, W[i].grad, L[i].grad), r, ("wq", "row_min", "row_max", "log_range", "g_w", "g_log_s")):
                assert torch.equal(a.detach(), b), (mode, i, shapes[i], nm)


def test_multi_tensor_launch_serves_the_layers_and_unused_outputs(fq):
    """prequantize_weights: one forward and one backward launch for a model's conv weights, same
    losses and gradients as layer-by-layer quantization (explicit per-shape noise off: LSQ)."""
    from mhaq_b200 import harness
    from mhaq_b200.quantization.gdnsq.layers._multi import prequantize_weights
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_conv2d import NoisyConv2d
    torch.manual_seed(0)
    x = torch.randn(32, 3, 32, 32, device="cuda")
    t = torch.randint(0, 10, (32,), device="cuda")
    q = harness.build_qat("resnet20", "cuda", qnmethod="LSQ", act_bit=4, weight_bit=4, distillation=False,
                          num_classes=10, calib_batch=x)
    q.train(); q.wrapped_criterion.train()
    convs = [m for m in q.model.modules() if isinstance(m, NoisyConv2d)]
    assert len(convs) == 18
    served = prequantize_weights(q.model)
    assert served == 18 and all(m.cache_probe()[1] is not None for m in convs)
    for m in convs:
        m._wq_cache.clear()
    grads = []
    # (cuDNN: no TF32, deterministic algorithms — otherwise two runs of the SAME graph already
    # differ by ~1e-4 of the largest gradient)
    saved_flags = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = False, True, False
    for use_multi in (True, False):
        import mhaq_b200.quantization.gdnsq.gdnsq_quant as GQ
        real = GQ.prequantize_weights
        GQ.prequantize_weights = real if use_multi else (lambda model: 0)
        try:
            fq.ops.set_device_philox_state(torch.tensor([5, 0], dtype=torch.int64, device="cuda"))
            fq.ops.reset_philox_call_counter()
            q.wrapped_criterion.loss_sum, q.wrapped_criterion.cnt = 0.0, 1
            loss = q.training_step((x, t), 0)
            loss.backward()
        finally:
            GQ.prequantize_weights = real
            fq.ops.set_device_philox_state(None)
        grads.append((loss.detach().clone(), {n: p.grad.clone() for n, p in q.named_parameters() if p.grad is not None}))
        q.zero_grad(set_to_none=True)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = saved_flags
    (l1, g1), (l2, g2) = grads
    assert torch.equal(l1, l2)
    assert set(g1) == set(g2)
    for n in g1:
        if "log_act_s" in n:      # the activations' noise streams are numbered differently in the two runs
            continue
        H.assert_close_rel(g1[n], g2[n], 1e-5, n, abs_floor=1e-6 * float(g2[n].abs().max()) + 1e-12)


@pytest.mark.parametrize("method,distill", [("STE", True), ("LSQ", False), ("AEWGS", True)])
def test_gradient_funnel_gives_the_same_gradients_with_fewer_launches(fq, method, distill):
    """quantization/gdnsq/_funnel.py: PotentialLoss's gradients w.r.t. log_act_s / log_act_q /
    log_wght_s return through the quantizer nodes and are added inside their backward kernels.
    Same loss, bitwise the same gradient for EVERY parameter (an fp32 add commutes; same Philox
    streams in both runs), and 2 launches per activation quantizer + 1 per conv weight fewer."""
    import contextlib
    from torch.profiler import profile, ProfilerActivity
    from mhaq_b200 import harness
    import mhaq_b200.quantization.gdnsq._funnel as FN
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_act import NoisyAct
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_conv2d import NoisyConv2d
    torch.manual_seed(0)
    x = torch.randn(32, 3, 32, 32, device="cuda")
    t = torch.randint(0, 10, (32,), device="cuda")
    q = harness.build_qat("resnet20", "cuda", qnmethod=method, act_bit=4, weight_bit=4, distillation=distill,
                          num_classes=10, calib_batch=x)
    q.train(); q.wrapped_criterion.train()
    if getattr(q, "tmodel", None) is not None:
        q.tmodel.eval()
    n_act = sum(isinstance(m, NoisyAct) for m in q.model.modules())
    n_conv = sum(isinstance(m, NoisyConv2d) for m in q.model.modules())
    saved_flags = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = False, True, False
    real_step = FN.step
    runs = []
    try:
        for funnel in (True, False):
            FN.step = real_step if funnel else contextlib.nullcontext

            def one_step():
                fq.ops.set_device_philox_state(torch.tensor([5, 0], dtype=torch.int64, device="cuda"))
                fq.ops.reset_philox_call_counter()
                q.wrapped_criterion.loss_sum, q.wrapped_criterion.cnt = 0.0, 1
                loss = q.training_step((x, t), 0)
                loss.backward()
                return loss
            one_step(); q.zero_grad(set_to_none=True)                  # warm-up (cuDNN plans, arenas)
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                loss = one_step()
                torch.cuda.synchronize()
            launches = sum(ev.count for ev in prof.key_averages()
                           if (getattr(ev, "device_time_total", 0.0) or 0.0) > 0)
            runs.append((loss.detach().clone(), {n: p.grad.clone() for n, p in q.named_parameters() if p.grad is not None},
                         launches))
            q.zero_grad(set_to_none=True)
    finally:
        FN.step = real_step
        fq.ops.set_device_philox_state(None)
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = saved_flags
    (l1, g1, k1), (l2, g2, k2) = runs
    assert torch.equal(l1, l2)
    assert set(g1) == set(g2)
    for n in g1:
        assert torch.equal(g1[n], g2[n]), n
    assert not FN._holders and not any("_funnel_act" in m.__dict__ or "_funnel_ls" in m.__dict__ for m in q.model.modules())
    assert k2 - k1 >= 2 * n_act + n_conv - 2, (k1, k2, n_act, n_conv)


# ---------------------------------------------------------------------------
# fused PotentialLoss arithmetic (mhaq_fq_potential_loss_*)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("lossless", [False, True])
@pytest.mark.parametrize("n_w,n_a", [(3904, 16), (7, 3), (1, 1)])
def test_fused_potential_loss_matches_the_torch_expression(fq, n_w, n_a, lossless):
    from mhaq_b200.quantization.gdnsq.gdnsq_loss import PotentialLossNoPred
    g = torch.Generator().manual_seed(n_w + n_a)
    lws = (torch.randn(n_w, generator=g) - 6).cuda()
    lwq = (lws.cpu() + 4 + torch.randn(n_w, generator=g)).cuda()        # around the 4-bit target: some active
    las = (torch.randn(n_a, generator=g) - 3).cuda()
    laq = (las.cpu() + 4 + 0.5 * torch.randn(n_a, generator=g)).cuda()
    if n_w > 2:
        lwq[1] = lws[1] + (4 - 1e-3)                                   # near / at the threshold
    base = torch.tensor(1.7, device="cuda")

    def run(fused):
        crit = PotentialLossNoPred(None, p=1, a=4, w=4, lossless=lossless)
        crit.t, crit.loss_sum, crit.cnt = 0.35, 2.5, 3
        crit.train()
        if not fused:
            crit._fusable = lambda *a: False
        leaves = [v.clone().requires_grad_(True) for v in (base, las, laq, lws, lwq)]
        out = []
        for step in range(2):                                          # second step: updated calibration state
            for v in leaves:
                v.grad = None
            loss = crit((leaves[0] * 1.0, *leaves[1:]))
            loss.backward()
            out.append((loss.detach().clone(), [v.grad.clone() for v in leaves],
                        [getattr(crit, k).detach().clone().float() for k in
                         ("wloss", "aloss", "rloss", "s_weight_loss", "q_weight_loss", "s_act_loss", "q_act_loss",
                          "weight_reg_loss")],
                        float(crit.loss_sum), float(crit.cnt)))
        return out

    for (l_f, g_f, logs_f, ls_f, c_f), (l_t, g_t, logs_t, ls_t, c_t) in zip(run(True), run(False)):
        H.assert_close_rel(l_f, l_t, 1e-6, "ploss", abs_floor=1e-7)
        for a, b, nm in zip(g_f, g_t, ("g_base", "g_las", "g_laq", "g_lws", "g_lwq")):
            H.assert_close_rel(a, b, 1e-6, nm, abs_floor=1e-9)
        for a, b in zip(logs_f, logs_t):
            H.assert_close_rel(a, b, 1e-6, "logged term", abs_floor=1e-7)
        assert abs(ls_f - ls_t) <= 1e-6 * abs(ls_t) and c_f == c_t


@pytest.mark.parametrize("n_tensors", [5, 30])
def test_multi_tensor_aewgs_equals_the_streaming_aewgs_path(fq, n_tensors):
    """AEWGS weights, model-wide: statistics kernel -> ONE packed all-reduce (a no-op on one rank)
    -> apply kernel, against the per-layer streaming path (row statistics, forward, AEWGS statistics
    + finalize, backward + finalize, amin scatter) that the goldens and the live-reference tests pin."""
    ops = fq.ops
    g = torch.Generator().manual_seed(100 + n_tensors)
    shapes = [MULTI_SHAPES[i % len(MULTI_SHAPES)] for i in range(n_tensors)]
    ws = [(torch.randn(s, generator=g) * 0.2).cuda() for s in shapes]
    lss = [(torch.full((s[0],) + (1,) * (len(s) - 1), -3.0) + 0.5 * torch.rand((s[0],) + (1,) * (len(s) - 1), generator=g)).cuda()
           for s in shapes]
    gos = [torch.randn(s, generator=g).cuda() for s in shapes]
    gls = [torch.randn(s[0], generator=g).cuda() for s in shapes]
    noises = [(torch.randint(0, 2, s, generator=g).float() - 0.5).cuda() for s in shapes]
    ref = []
    for i in range(n_tensors):
        w, ls = ws[i].clone().requires_grad_(True), lss[i].clone().requires_grad_(True)
        wq, mn, mx = ops.weight_fake_quant_log(w, ls, method="AEWGS", noise=noises[i])
        lr = torch.log2(mx - mn + torch.exp2(ls.ravel()))
        ((wq * gos[i]).sum() + (lr * gls[i]).sum()).backward()
        ref.append((wq.detach(), lr.detach(), w.grad, ls.grad))
    W = [w.clone().requires_grad_(True) for w in ws]
    L = [l.clone().requires_grad_(True) for l in lss]
    res = ops.weight_fake_quant_rows_multi(W, L, method="AEWGS", noises=noises)
    loss = sum((wq * go).sum() + (lr * gl).sum() for (wq, mn, mx, lr), go, gl in zip(res, gos, gls))
    loss.backward()
    for i, ((wq, mn, mx, lr), (wq_r, lr_r, gw_r, gls_r)) in enumerate(zip(res, ref)):
        assert torch.equal(wq.detach(), wq_r), (i, shapes[i])
        assert torch.equal(lr.detach(), lr_r), (i, shapes[i])
        n_inner = ws[i][0].numel()
        # same per-element arithmetic; the per-channel means differ in summation order (1e-5 on gx),
        # the zero-point gradient that lands on the row minimum is a difference of two O(N) sums
        flat = ws[i].reshape(shapes[i][0], -1)
        ties = ((flat == flat.amin(1, keepdim=True)) | (flat == flat.amax(1, keepdim=True))).reshape(shapes[i]).cpu()
        a, b = W[i].grad.cpu(), gw_r.cpu()
        H.assert_close_rel(a[~ties], b[~ties], REL, f"g_w[{i}]", abs_floor=1e-7)
        H.assert_close_rel(a[ties], b[ties], REL, f"g_w[{i}] (row extrema)", abs_floor=max(2e-5, 1.5e-6 * math.sqrt(n_inner)))
        H.assert_close_rel(L[i].grad, gls_r, REL, f"g_log_s[{i}]", abs_floor=4e-7 * math.sqrt(n_inner) + 1e-7)
