"""Validation-time bit-width statistics — the public functions of the reference's
src/quantization/gdnsq/utils/model_stats.py:116-262, computed with tensor ops on the device
(the reference loops over channels in Python with an ``.item()`` sync per channel).
SURVEY.md §8 row (f)-3: not on the training hot path."""
import torch

from ....aux.types import QScheme, is_per_channel, is_per_tensor
from ..layers.gdnsq_act import NoisyAct
from ..layers.gdnsq_conv2d import NoisyConv2d
from ..layers.gdnsq_linear import NoisyLinear

_WEIGHT_LAYERS = (NoisyConv2d, NoisyLinear)


def get_activations_bit_width(log_q, log_s, b):
    return (log_q - log_s).mean()


def _code_span_bits(module) -> torch.Tensor:
    """log2(#distinct code levels spanned) per channel (per-channel) or for the tensor."""
    with torch.no_grad():
        if hasattr(module, "quantized_weight"):
            module.quantized_weight()            # refresh Q.scale / Q.zero_point
        codes = module.Q.quantize(module.weight.detach())
        if is_per_channel(module.qscheme):
            flat = codes.reshape(codes.shape[0], -1)
            return torch.log2(flat.amax(1) - flat.amin(1) + 1)
        return torch.log2(codes.max() - codes.min() + 1).reshape(1)


def get_true_layer_bit_width(module, max=True):
    bits = _code_span_bits(module)
    return bits.max() if (max or is_per_tensor(module.qscheme)) else bits.mean()


def get_true_weights_width(model, max=True):
    widths = torch.stack([get_true_layer_bit_width(m) for m in model.modules()
                          if isinstance(m, _WEIGHT_LAYERS)])
    return widths.max() if max else widths.mean()


def get_true_activations_width(model, max=True):
    widths = torch.stack([m.bw.detach().float().reshape(()).to(next(model.parameters()).device)
                          for m in model.modules() if isinstance(m, NoisyAct)])
    return widths.max() if max else widths.mean()


def get_layer_wnb_bit_width(layer_weights, log_s, config=QScheme.PER_TENSOR):
    if is_per_tensor(config):
        mn, mx = layer_weights.amin(), layer_weights.amax()
    else:
        dims = tuple(range(1, layer_weights.dim()))
        mn, mx = layer_weights.amin(dims), layer_weights.amax(dims)
    log_q = torch.log2((mx - mn).reshape(log_s.shape) + torch.exp2(log_s))
    return get_activations_bit_width(log_q, log_s, 0)


def get_weights_bit_width_mean(model):
    vals = []
    for m in model.modules():
        if isinstance(m, _WEIGHT_LAYERS):
            bw = get_layer_wnb_bit_width(m.weight.detach(), m.log_wght_s.detach(), m.qscheme)
            vals.append(bw.mean())
    vals = torch.stack(vals)
    return vals[~torch.isnan(vals)].mean()


def get_activations_bit_width_mean(model):
    return torch.stack([get_activations_bit_width(m.log_act_q.detach(), m.log_act_s.detach(),
                                                  m.act_b.detach())
                        for m in model.modules() if isinstance(m, NoisyAct)]).mean()


def is_converged(model):
    loss = model.wrapped_criterion
    return bool(get_true_weights_width(model) <= loss.wt) and bool(
        get_true_activations_width(model) <= loss.at)


# name -> fn(model) logged by the patched validation step (gdnsq_quant.py:260-301)
VALIDATION_STATS = (
    ("Mean weights bit width", get_weights_bit_width_mean),
    ("Actual weights bit width", lambda m: get_true_weights_width(m, max=False)),
    ("Actual weights max bit width", get_true_weights_width),
    ("Mean activations bit width", get_activations_bit_width_mean),
    ("Actual activations bit widths", lambda m: get_true_activations_width(m, max=False)),
    ("Actual activations max bit widths", get_true_activations_width),
)
