"""``ModelHelper.get_model_values`` — mirror of the reference's
src/quantization/gdnsq/utils/model_helper.py:11-76: collects, per quantized layer, the
log-domain scale parameters and the differentiable weight range
``log2(max - min + 2^log_wght_s)`` that ``PotentialLoss`` constrains."""
import torch
from torch import nn

from ....aux.types import QScheme, is_per_channel, is_per_tensor
from ..layers.gdnsq_act import NoisyAct
from ..layers.gdnsq_conv2d import NoisyConv2d
from ..layers.gdnsq_linear import NoisyLinear


class ModelHelper:
    @staticmethod
    def get_model_values(model: nn.Module, qscheme: QScheme = QScheme.PER_TENSOR):
        log_wght_s, log_w_n_b, log_act_q, log_act_s = [], [], [], []
        for _, m in model.named_modules():
            if isinstance(m, (NoisyConv2d, NoisyLinear)):
                if not m.log_wght_s.requires_grad:
                    continue
                if is_per_channel(qscheme):
                    dims = tuple(range(1, m.weight.dim()))
                    # (this step's alias of log_wght_s when the gradient funnel is active: the
                    # loss's gradient then returns through the weight kernel, ../_funnel.py)
                    ls = m.__dict__.pop("_funnel_ls", None)
                    log_wght_s.append((m.log_wght_s if ls is None else ls).ravel())
                    # the layer already reduced the weight rows in this step's forward: the
                    # fused row kernels hand back log2(max - min + 2^log_wght_s) itself, the
                    # streaming path its differentiable (min, max) — either way no further
                    # passes over the weight
                    on_gpu = m.weight.is_cuda
                    lr = m.log_w_range() if on_gpu and hasattr(m, "log_w_range") else None
                    if lr is not None:
                        log_w_n_b.append(lr)
                        continue
                    rr = m.row_range() if on_gpu and hasattr(m, "row_range") else None
                    mn, mx = rr if rr is not None else (m.weight.amin(dims), m.weight.amax(dims))
                else:
                    log_wght_s.append(m.log_wght_s)
                    mn, mx = m.weight.amin(), m.weight.amax()
                # one LSB of head-room against overflow (model_helper.py:43-44)
                log_w_n_b.append(torch.log2(mx - mn + torch.exp2(m.log_wght_s.ravel())))
            elif isinstance(m, NoisyAct):
                if m.log_act_s.requires_grad:
                    las, laq = m.__dict__.pop("_funnel_act", None) or (m.log_act_s, m.log_act_q)
                    log_act_q.append(laq)
                    log_act_s.append(las)
        if is_per_tensor(qscheme):
            return (torch.stack(log_act_s).ravel(), torch.stack(log_act_q).ravel(),
                    torch.stack(log_wght_s).ravel(), torch.stack(log_w_n_b).ravel())
        return (torch.cat(log_act_s), torch.cat(log_act_q), torch.cat(log_wght_s),
                torch.cat(log_w_n_b))
