"""``NoisyAct`` — activation quantizer wrapper; mirror of the reference's
src/quantization/gdnsq/layers/gdnsq_act.py:9-55 on the sm_100a kernels.

Same constructor arguments, parameters (``log_act_s, log_act_q, act_b`` each
``(1,)``; ``act_b`` trainable iff ``signed``), initial values and ``bw`` attribute.
The parameter preparation (``exp2``, ``act_b + q - s``) stays in PyTorch so autograd
chains the log-domain gradients exactly as in the reference; the tensor-sized work
is one fused kernel each way.
"""
import torch
from torch import nn, inf

from .... import ops
from .. import _funnel
from ..gdnsq import Quantizer
from ..gdnsq_utils import QNMethod


class NoisyAct(nn.Module):
    def __init__(self, init_s=-10, init_q=10, signed=True, noise_ratio=1, disable=False,
                 qnmethod: QNMethod = QNMethod.STE) -> None:
        super().__init__()
        self.disable = disable
        self.signed = signed
        zero_point = 0.0 if not signed else -torch.exp2(torch.tensor(init_q - 1).float())
        self._act_b = torch.tensor([zero_point]).float()
        self._log_act_s = torch.tensor([init_s]).float()
        self._log_act_q = torch.tensor([init_q]).float()
        self._noise_ratio = torch.tensor(noise_ratio)
        self.log_act_q = torch.nn.Parameter(self._log_act_q, requires_grad=True)
        self.act_b = torch.nn.Parameter(self._act_b, requires_grad=bool(signed))
        self.log_act_s = torch.nn.Parameter(self._log_act_s, requires_grad=True)
        self.Q = Quantizer(self, torch.exp2(self._log_act_s), 0, -inf, inf, qnmethod=qnmethod)
        self.bw = torch.tensor(0.0)

    def _operands(self):
        """(scale, zero_point, min_val, max_val) exactly as the reference assigns them
        (gdnsq_act.py:42-48)."""
        s = torch.exp2(self.log_act_s)
        q = torch.exp2(self.log_act_q)
        return s, self.act_b, self.act_b, self.act_b + q - s

    def forward(self, x):
        if self.disable:
            return x
        self._in_minmax = None
        if (self.training and self.Q.positive_scale and x.is_cuda
                and self.Q._method().name != "AEWGS"):
            # (AEWGS on a per-tensor quantizer follows the reference's dim-0 statistics quirk,
            # which needs the linear operands: it takes the generic route below)
            # fused: the kernels read log_act_s / log_act_q / act_b themselves and return the
            # log-domain gradients; Q.scale & co. are materialised only if somebody reads them
            self.Q.defer(self._operands)
            if _funnel.active() and torch.is_grad_enabled() and self.log_act_s.requires_grad:
                # one gradient per parameter: PotentialLoss reads these aliases (via ModelHelper)
                # and its gradient comes back into this node's backward kernel (_funnel.py)
                y, las, laq = ops.act_fake_quant(x, self.log_act_s, self.log_act_q, self.act_b,
                                                 method=self.Q._method(), funnel=True)
                self._funnel_act = (las, laq)
                _funnel.hold(self)
                return y
            return ops.act_fake_quant(x, self.log_act_s, self.log_act_q, self.act_b,
                                      method=self.Q._method())
        self.Q.scale, self.Q.zero_point, self.Q.min_val, self.Q.max_val = self._operands()
        if self.training:
            return self.Q.fake_quant(x)
        # eval: the reference quantizes, asserts (3 host syncs), takes aminmax of the codes and
        # dequantizes (gdnsq_act.py:50-55); here one pass yields y and (min, max, #bad) codes
        needs_graph = torch.is_grad_enabled() and (
            x.requires_grad or any(p.requires_grad for p in (self.log_act_s, self.log_act_q, self.act_b)))
        with torch.no_grad():
            y, mm = self.Q.fake_quant_eval(x)
        self.Q._assert_valid(mm)
        self.bw = torch.log2(mm[1] - mm[0] + 1)
        # the same pass also took min / max of the INPUT (mm[3], mm[4]): what a MinMaxObserver
        # forward hook wants from this very tensor during calibration (calib/minmaxobserver.py)
        self._in_minmax = (x.data_ptr(), x._version, tuple(x.shape), mm[3:5])
        if needs_graph:
            return self.Q.fake_quant(x)
        return y
