"""Operator boundary of the GDNSQ quantizer — mirror of the reference's
src/quantization/gdnsq/gdnsq.py:159-241 (class ``Quantizer``), backed by the
sm_100a kernels.

Same constructor, same mutable attributes (``scale, zero_point, min_val, max_val,
qnmethod, module, rnoise_ratio, positive_scale``), same ``quantize`` /
``dequantize`` semantics and eval-mode assertion messages.  New: ``fake_quant``,
the fused ``dequantize(quantize(x))`` the layer wrappers call (one kernel each
way instead of ~8 / ~19-27 ATen launches).

The reference's ``QNoise`` / ``QNSTE`` / ``QNLSQ`` / ``QNEWGS`` / ``QNAEWGS``
autograd Functions (gdnsq.py:11-147) are fused into ``mhaq_fq_fwd_f32`` /
``mhaq_fq_bwd_f32``; the names stay importable and ``apply`` keeps working through
``ops.rounding_noise`` (the two-step ``v + QN*.apply(v, s)`` form).
"""
from __future__ import annotations

from enum import Enum

import torch
from torch import Tensor

from ... import ops
from .gdnsq_utils import QMode, QNMethod  # noqa: F401


def reduce_to_shape(t: Tensor, like: Tensor) -> Tensor:
    """Mean over the dims where `like` has size 1 (reference gdnsq.py:150-152)."""
    dims = tuple(i for i, n in enumerate(like.shape) if n == 1)
    return torch.mean(t, dim=dims, keepdim=True)


class QNoise:
    """The reference's rounding-noise autograd Functions (gdnsq.py:11-147): ``apply(v, s)`` returns
    ``round(v) - v`` and its backward is the estimator's (``grad_v``, ``grad_s``).  In this package
    they are fused into the kernels; ``apply`` is kept working for callers of the two-step form
    (``v + QN*.apply(v, s)``, gdnsq.py:206-208, and ``scaled_noise``): it runs the forward kernel on
    the already-scaled value (zero point 0, unit divisor) and returns ``codes - v``, so the codes are
    bit-identical and autograd sees d(noise)/dv = estimator - 1, d(noise)/ds = the estimator's
    scale gradient, exactly the reference's (backward kernel with the code gradient as input).
    ``QNoise.apply`` itself raises in backward in the reference (gdnsq.py:26-29); here it behaves
    like QNSTE."""
    _method = "STE"

    @classmethod
    def apply(cls, value, scale):
        return ops.rounding_noise(value, scale, method=cls._method)


class QNSTE(QNoise):
    _method = "STE"


class QNLSQ(QNoise):
    _method = "LSQ"


class QNEWGS(QNoise):
    _method = "EWGS"


class QNAEWGS(QNoise):
    _method = "AEWGS"


def scaled_noise(x, s):
    return QNoise.apply(x, s)


class Quantizer:
    def __init__(self, module, scale, zero_point, min_val, max_val,
                 rnoise_ratio=torch.Tensor([-1.0]), qnmethod: QNMethod = QNMethod.STE) -> None:
        self._lazy = None
        self.module = module
        self.scale = scale
        self.zero_point = zero_point
        self.min_val = min_val
        self.max_val = max_val
        self.rnoise_ratio = torch.Tensor([rnoise_ratio])       # vestigial (gdnsq.py:185)
        # evaluated once at construction, exactly like the reference (gdnsq.py:186)
        self.positive_scale = torch.all(torch.as_tensor(self.scale) > 0).item()
        self.qnmethod = qnmethod

    # The four operands stay ordinary read/write attributes (the reference's layers assign
    # them before every call, gdnsq_act.py:45-48).  The fused layer paths hand the kernels the
    # log-domain parameters instead and only *defer* these assignments: `defer(fn)` installs a
    # thunk that produces (scale, zero_point, min_val, max_val) — with autograd history, exactly
    # what the reference would have assigned — the first time any of them is read.
    def defer(self, fn) -> None:
        self._lazy = fn

    def _resolve(self):
        fn = self._lazy
        if fn is not None:
            self._lazy = None
            self._scale, self._zero_point, self._min_val, self._max_val = fn()

    def _get(name):
        def getter(self):
            self._resolve()
            return getattr(self, "_" + name)

        def setter(self, value):
            self._resolve()
            setattr(self, "_" + name, value)

        return property(getter, setter)

    scale = _get("scale")
    zero_point = _get("zero_point")
    min_val = _get("min_val")
    max_val = _get("max_val")
    del _get

    # -- helpers ---------------------------------------------------------------------------
    def _method(self):
        """The estimator as this package's QNMethod.  Another package's enum with the same
        name AND value is accepted — the reference's own `QNMethod` when its `GDNSQQuant`
        constructs these layer classes; anything else raises like
        gdnsq.py:240-241."""
        m = self.qnmethod
        if isinstance(m, QNMethod):
            return m
        if isinstance(m, Enum) and m.name in QNMethod.__members__ and QNMethod[m.name].value == m.value:
            return QNMethod[m.name]
        raise AttributeError(f"Unknown method {self.qnmethod}!")

    def _zp_like(self, value):
        zp = self.zero_point
        if not torch.is_tensor(zp):
            zp = torch.full((1,), float(zp), dtype=value.dtype, device=value.device)
        return zp

    # -- reference API ---------------------------------------------------------------------
    def quantize(self, value):
        """Clamp, shift, scale, round: integer-valued codes (gdnsq.py:189-219)."""
        if not self.positive_scale:                            # gdnsq.py:201-202
            return torch.clamp(value, min=self.min_val, max=self.max_val) - self.zero_point
        method = self._method()
        training = self.module.training
        if training or torch.is_grad_enabled() and value.requires_grad:
            codes = ops.quantize_codes(value, self.scale, self._zp_like(value), self.min_val,
                                       self.max_val, method=method)
            if training:
                return codes
            self._assert_valid(ops.quantize_eval(value, self.scale, self._zp_like(value),
                                                 self.min_val, self.max_val, want_y=False)[2])
            return codes
        _, codes, mm = ops.quantize_eval(value, self.scale, self._zp_like(value), self.min_val,
                                         self.max_val, want_y=False, want_codes=True)
        self._last_minmax = mm
        self._assert_valid(mm)
        return codes

    @staticmethod
    def _assert_valid(mm):
        """Eval-mode assertions of gdnsq.py:211-217.  Codes are rint() of a clamped value, so
        they are integral and inside [floor((min-zp)/s), ceil((max-zp)/s)] by construction
        unless non-finite; the kernel counts non-finite codes in the same pass (one host
        sync instead of the reference's three)."""
        if mm is not None and float(mm[2].item()) != 0.0:
            raise AssertionError("Not all elements in the tensor have integer values.")

    def dequantize(self, quantized_value):
        """codes * scale + zero_point (gdnsq.py:221-229) — the two-call form; the layers use
        fake_quant()."""
        if not self.positive_scale:
            return quantized_value + self.zero_point
        return quantized_value * self.scale + self.zero_point

    def _get_rnoise(self, value: Tensor, scale: Tensor):
        return {"STE": QNSTE, "EWGS": QNEWGS, "AEWGS": QNAEWGS, "LSQ": QNLSQ}[self._method().name].apply(value, scale)

    # -- fused entry -----------------------------------------------------------------------
    def fake_quant(self, value, noise=None):
        """dequantize(quantize(value)) in one kernel each way."""
        if not self.positive_scale:
            return torch.clamp(value, min=self.min_val, max=self.max_val)
        return ops.fake_quant(value, self.scale, self._zp_like(value), self.min_val, self.max_val,
                              method=self._method(), noise=noise)

    def fake_quant_weight(self, weight, log_scale=None, noise=None):
        """Per-channel weight path with the row-minimum zero point computed in the same pass
        (sets self.zero_point like NoisyConv2d.forward does, gdnsq_conv2d.py:80-84).
        With `log_scale` (= log_wght_s) the kernels take the scale in the log domain and
        return d/d log_wght_s directly; self.scale is then materialised lazily.  Short rows
        (conv / linear weights) and a method other than AEWGS take the row-resident fused
        kernels: one launch each way, which also yield ModelHelper's
        log2(row_max - row_min + 2^log_wght_s) (utils/model_helper.py:24-25,44).
        Returns (weight_q, row_min, row_max, log_range or None)."""
        pshape = (weight.shape[0],) + (1,) * (weight.dim() - 1)
        lr = None
        if log_scale is not None:
            if ops.weight_rows_fusable(weight, log_scale, self._method()):
                wq, mn, mx, lr = ops.weight_fake_quant_rows(weight, log_scale, method=self._method(),
                                                            noise=noise)
            else:
                wq, mn, mx = ops.weight_fake_quant_log(weight, log_scale, method=self._method(), noise=noise)
            zp = mn.view(pshape)
            self.defer(lambda: (torch.exp2(log_scale).reshape(pshape), zp, self._min_val, self._max_val))
        else:
            wq, mn, mx = ops.weight_fake_quant(weight, self.scale, method=self._method(), noise=noise)
            self.zero_point = mn.view(pshape)
        return wq, mn, mx, lr

    def fake_quant_eval(self, value):
        """No-grad fused forward that also yields (min code, max code, #non-finite) — the
        eval extras of gdnsq.py:211-217 / gdnsq_act.py:51-54 without extra passes."""
        y, _, mm = ops.quantize_eval(value, self.scale, self._zp_like(value), self.min_val,
                                     self.max_val, want_y=True, want_codes=False)
        return y, mm
