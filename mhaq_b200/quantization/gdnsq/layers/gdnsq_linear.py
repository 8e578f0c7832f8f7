"""``NoisyLinear`` — weight-quantizing linear layer; mirror of the reference's
src/quantization/gdnsq/layers/gdnsq_linear.py:13-95 on the sm_100a kernels.

State-dict layout as in the reference: ``log_wght_s`` is ``(1,)`` (per-tensor) or
``(out_features,1,1,1)`` (per-channel, gdnsq_linear.py:54-58).  The reference's
per-channel forward calls ``weight.amin((1,2,3))`` on a 2-D weight and therefore
cannot run (gdnsq_linear.py:71); here the per-channel scale is
viewed as ``(out_features, 1)`` and the row minimum taken over dim 1 — the evident
intent, same numbers the conv path produces for a 1x1 kernel.
"""
import torch
import torch.nn.functional as F
from torch import nn, inf

from ....aux.types import QScheme, is_per_channel, is_per_tensor
from ....aux.qutils import is_biased
from ..gdnsq import Quantizer
from ..gdnsq_utils import QNMethod
from ._wcache import WeightQuantCache


class NoisyLinear(nn.Linear):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None,
                 dtype=None, qscheme: QScheme = QScheme.PER_TENSOR, log_s_init: float = -12,
                 rand_noise: bool = False, qnmethod: QNMethod = QNMethod.STE) -> None:
        super().__init__(in_features, out_features, bias, device, dtype)
        self.qscheme = qscheme
        if is_per_tensor(self.qscheme):
            self.log_wght_s = nn.Parameter(torch.Tensor([log_s_init]), requires_grad=True)
        elif is_per_channel(self.qscheme):
            self.log_wght_s = nn.Parameter(torch.empty((out_features, 1, 1, 1)).fill_(log_s_init),
                                           requires_grad=True)
        self._noise_ratio = nn.Parameter(torch.Tensor([1]), requires_grad=False)
        self.Q = Quantizer(self, torch.exp2(self.log_wght_s), 0, -inf, inf, qnmethod=qnmethod)
        self.rand_noise = rand_noise
        self._wq_cache = WeightQuantCache()

    def quantized_weight(self):
        return self._quantize()[0]

    def row_range(self):
        """(row_min, row_max), differentiable, from the quantization pass; None per-tensor."""
        out = self._quantize()
        return (out[1], out[2]) if out[1] is not None else None

    def log_w_range(self):
        """log2(row_max - row_min + 2^log_wght_s) from the fused row kernels, or None."""
        return self._quantize()[3]

    def _quantize(self):
        key, hit = self._wq_cache.lookup((self.weight, self.log_wght_s), torch.is_grad_enabled(),
                                         self.training)
        if hit is not None:
            return hit
        if is_per_channel(self.qscheme):
            if self.Q.positive_scale and self.weight.is_cuda:
                out = self.Q.fake_quant_weight(self.weight, log_scale=self.log_wght_s)
            else:
                self.Q.scale = torch.exp2(self.log_wght_s).reshape(self.out_features, 1)
                self.Q.zero_point = self.weight.amin(1, keepdim=True)
                out = (self.Q.fake_quant(self.weight), None, None, None)
        else:
            self.Q.scale = torch.exp2(self.log_wght_s)
            self.Q.zero_point = self.weight.amin()
            out = (self.Q.fake_quant(self.weight), None, None, None)
        self._wq_cache.store(key, out)
        return out

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return F.linear(input, self.quantized_weight(), self.bias)

    def extra_repr(self) -> str:
        noise_ratio = self._noise_ratio if self.rand_noise else torch.zeros_like(self._noise_ratio)
        return (f"in_features={self.in_features}, out_features={self.out_features}, "
                f"bias={is_biased(self)},\nlog_wght_s={self.log_wght_s}, noise_ratio={noise_ratio}")
