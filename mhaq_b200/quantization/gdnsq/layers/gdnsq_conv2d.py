"""``NoisyConv2d`` — weight-quantizing convolution; mirror of the reference's
src/quantization/gdnsq/layers/gdnsq_conv2d.py:13-119 on the sm_100a kernels.

Same constructor, parameters and state-dict layout: ``log_wght_s`` is ``(1,)``
(per-tensor) or ``(O,1,1,1)`` (per-channel); ``log_b_s`` ``(1,)`` exists only
per-channel and — exactly as in the reference — is never used (bias quantization
reuses the weight scale / row minimum, gdnsq_conv2d.py:86-94); ``_noise_ratio`` is
a non-trainable ``(1,)`` Parameter kept for state-dict compatibility.
"""
import torch
from torch import nn, inf

from ....aux.types import QScheme, is_per_channel, is_per_tensor
from ....aux.qutils import is_biased
from ..gdnsq import Quantizer
from ..gdnsq_utils import QNMethod
from ._wcache import WeightQuantCache


class NoisyConv2d(nn.Conv2d):
    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride=1, padding=0,
                 dilation=1, groups: int = 1, bias: bool = True, padding_mode: str = "zeros",
                 device=None, dtype=None, qscheme: QScheme = QScheme.PER_TENSOR,
                 log_s_init: float = -12, rand_noise: bool = False, quant_bias: bool = False,
                 qnmethod: QNMethod = QNMethod.AEWGS) -> None:
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         bias, padding_mode, device, dtype)
        self.qscheme = qscheme
        if is_per_tensor(self.qscheme):
            self.log_wght_s = nn.Parameter(torch.Tensor([log_s_init]), requires_grad=True)
        elif is_per_channel(self.qscheme):
            self.log_wght_s = nn.Parameter(torch.empty((out_channels, 1, 1, 1)).fill_(log_s_init),
                                           requires_grad=True)
            self.log_b_s = nn.Parameter(torch.empty(1).fill_(log_s_init), requires_grad=True)
        self._noise_ratio = torch.nn.Parameter(torch.Tensor([1]), requires_grad=False)
        self.Q = Quantizer(self, torch.exp2(self.log_wght_s), 0, -inf, inf, qnmethod=qnmethod)
        self.rand_noise = rand_noise
        self.quant_bias = quant_bias
        if self.quant_bias:
            self.Q_b = Quantizer(self, torch.exp2(self.log_b_s), 0, -inf, inf, qnmethod=qnmethod)
        self._wq_cache = WeightQuantCache()

    def quantized_weight(self):
        """(weight_q, bias_q) — computed once per parameter version (see _wcache)."""
        out = self._quantize()
        return out[0], out[1]

    def row_range(self):
        """(row_min, row_max) of the weight, differentiable, from the same pass that produced
        the quantized weight (consumed by ModelHelper.get_model_values); None per-tensor."""
        out = self._quantize()
        return (out[2], out[3]) if out[2] is not None else None

    def log_w_range(self):
        """log2(row_max - row_min + 2^log_wght_s) per output channel, differentiable, when the
        fused row kernels produced it in this step's quantization pass (else None)."""
        return self._quantize()[4]

    def _quantize(self):
        key, hit = self._wq_cache.lookup((self.weight, self.log_wght_s, self.bias),
                                         torch.is_grad_enabled(), self.training)
        if hit is not None:
            return hit
        mx = lr = None
        if is_per_channel(self.qscheme) and self.positive_scale_ok():
            # fused: row min (zero point) + row max in one pass, quantization in the next, the
            # scale taken in the log domain (no exp2 / Exp2Backward launches)
            weight, mn_flat, mx, lr = self.Q.fake_quant_weight(self.weight, log_scale=self.log_wght_s)
            s = mn = None
        else:
            s = torch.exp2(self.log_wght_s)
            self.Q.scale = s
            if is_per_channel(self.qscheme):
                mn = self.weight.amin((1, 2, 3), keepdim=True)
            else:
                mn = self.weight.amin()
            self.Q.zero_point = mn
            mn_flat = None
            weight = self.Q.fake_quant(self.weight)

        if self.quant_bias:
            if s is None:
                s, mn = self.Q.scale, self.Q.zero_point
            self.Q_b.scale = s.ravel()
            self.Q_b.zero_point = mn.ravel()
            bias = self.Q_b.fake_quant(self.bias)
        else:
            bias = self.bias
        out = (weight, bias, mn_flat, mx, lr)
        # a quantized bias (tiny) is recomputed per call so that one cache entry never pins
        # two autograd graphs
        if not self.quant_bias:
            self._wq_cache.store(key, out)
        return out

    # ---- multi-tensor launch support (layers/_multi.py) --------------------------------------
    def multi_ok(self) -> bool:
        """May this layer's weight be quantized by the model-wide multi-tensor launch?"""
        from .... import ops
        return (is_per_channel(self.qscheme) and not self.quant_bias and self.positive_scale_ok()
                and ops.weight_rows_fusable(self.weight, self.log_wght_s, self.Q._method(), multi=True))

    def cache_probe(self):
        """(key, hit) of this step's weight-cache entry."""
        return self._wq_cache.lookup((self.weight, self.log_wght_s, self.bias),
                                     torch.is_grad_enabled(), self.training)

    def adopt_quantized(self, key, wq, mn, mx, lr, ls_alias=None):
        """Install (weight_q, row_min, row_max, log_range) computed by the multi-tensor launch as
        this step's cache entry, exactly what `_quantize` would have stored.  `ls_alias`: this
        step's alias of log_wght_s for ModelHelper (gradient funnel, ../_funnel.py)."""
        if ls_alias is not None:
            from .. import _funnel
            self._funnel_ls = ls_alias
            _funnel.hold(self)
        pshape = (self.weight.shape[0],) + (1,) * (self.weight.dim() - 1)
        log_s, q = self.log_wght_s, self.Q
        q.defer(lambda: (torch.exp2(log_s).reshape(pshape), mn.view(pshape), q._min_val, q._max_val))
        self._wq_cache.store(key, (wq, self.bias, mn, mx, lr))

    def positive_scale_ok(self):
        return self.Q.positive_scale and self.weight.is_cuda

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        weight, bias = self.quantized_weight()
        return self._conv_forward(input, weight, bias)

    def extra_repr(self) -> str:
        noise_ratio = self._noise_ratio if self.rand_noise else torch.zeros_like(self._noise_ratio)
        return (f"in_channels={self.in_channels}, out_channels={self.out_channels}, "
                f"kernel_size={self.kernel_size},\nstride={self.stride}, padding={self.padding}, "
                f"dilation={self.dilation},\ngroups={self.groups}, bias={is_biased(self)}, "
                f"log_wght_s_mean={self.log_wght_s.mean()},\nnoise_ratio={noise_ratio}, "
                f"quantized_bias={self.quant_bias}")
