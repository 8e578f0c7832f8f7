"""Comparison rules shared by tests/ and __graft_entry__.smoke().  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import torch


def assert_bit_exact(a, b, what=""):
    """Numerically identical including NaN positions (signed zeros compare equal)."""
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    a, b = a.detach().cpu(), b.detach().cpu()
    same = (a == b) | (torch.isnan(a) & torch.isnan(b))
    if not bool(same.all()):
        bad = (~same).nonzero()
        i = tuple(bad[0].tolist())
        raise AssertionError(f"{what}: {int((~same).sum())}/{a.numel()} elements differ; "
                             f"first at {i}: {a[i].item()!r} vs {b[i].item()!r}")


def assert_close_rel(a, b, rel, what="", abs_floor=0.0):
    """|a-b| <= rel*|b| + abs_floor elementwise (b is the reference)."""
    a, b = a.detach().cpu().double().reshape(-1), b.detach().cpu().double().reshape(-1)
    assert a.shape == b.shape, f"{what}: shape mismatch"
    err = (a - b).abs()
    tol = rel * b.abs() + abs_floor
    if not bool((err <= tol).all()):
        i = int((err - tol).argmax())
        raise AssertionError(f"{what}: |{a[i].item():.9g} - {b[i].item():.9g}| = {err[i].item():.3g} "
                             f"> {tol[i].item():.3g} (rel {rel})")


def assert_param_grad(ours, ref32, exact, rel, what, abs_floor):
    """Parameter-gradient rule.

    A parameter gradient is a sum of N signed fp32 terms.  The reference accumulates it
    in fp32, in separately-rounded pieces that nearly cancel, so its own value moves by
    more than `rel` with the summation order once the sum is ill-conditioned.  Pass if
        |ours - ref32| <= rel*|ref32| + abs_floor                        (the stated bar), or
        |ours - exact| <= max(rel*|exact|, |ref32 - exact|) + abs_floor  (at least as close to
                                    the exact sum of the reference's terms as the reference is)
    `exact` = fp64 sum of the reference's own fp32 per-element terms."""
    ours = ours.detach().cpu().double().reshape(-1)
    ref32 = ref32.detach().cpu().double().reshape(-1)
    exact = exact.detach().cpu().double().reshape(-1)
    ok1 = (ours - ref32).abs() <= rel * ref32.abs() + abs_floor
    ok2 = (ours - exact).abs() <= torch.maximum(rel * exact.abs(), (ref32 - exact).abs()) + abs_floor
    ok = ok1 | ok2
    if not bool(ok.all()):
        i = int((~ok).nonzero()[0])
        den = max(abs(ref32[i].item()), 1e-30)
        raise AssertionError(
            f"{what}[{i}]: ours {ours[i].item():.9g} ref32 {ref32[i].item():.9g} exact {exact[i].item():.9g} "
            f"(|ours-ref32|/|ref32| = {abs(ours[i] - ref32[i]).item() / den:.2e}, "
            f"|ours-exact|/|ref32| = {abs(ours[i] - exact[i]).item() / den:.2e}, "
            f"|ref32-exact|/|ref32| = {abs(ref32[i] - exact[i]).item() / den:.2e})")


def exact_param_grads(fq, x, go, scale, zp, lo, hi, method, noise):
    """fp64 per-channel sums of the per-element parameter-gradient terms that `fq`
    (the oracle) produces when every parameter is expanded to a full-size leaf.
    Not valid for AEWGS (expanding the scale changes reduce_to_shape's dims)."""
    full = lambda t: None if t is None else t.detach().expand(x.shape).contiguous().requires_grad_(True)
    P = [full(p) if torch.is_tensor(p) else None for p in (scale, zp, lo, hi)]
    import math
    y = fq(x.detach(), P[0], P[1], -math.inf if P[2] is None else P[2],
           math.inf if P[3] is None else P[3], method=method, noise=noise)
    y.backward(go)
    out = []
    for pf, p in zip(P, (scale, zp, lo, hi)):
        if pf is None:
            out.append(None)
            continue
        g64 = pf.grad.double()
        if p.numel() == 1:
            out.append(g64.sum().reshape(p.shape).cpu())
        else:
            out.append(g64.reshape(p.numel(), -1).sum(1).reshape(p.shape).cpu())
    return out
