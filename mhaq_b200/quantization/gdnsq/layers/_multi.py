"""Model-wide weight fake-quantization: ONE launch forward, ONE backward for every per-channel
conv weight of a model (reference: one `Q.dequantize(Q.quantize(w))` chain
per layer per forward, gdnsq_conv2d.py:98, plus ModelHelper's range term per layer,
model_helper.py:24-25,44).

`prequantize_weights(model)` is called at the top of the patched training step.  It quantizes the
weights of all eligible `NoisyConv2d` layers with `ops.weight_fake_quant_rows_multi` and installs
the results as the layers' per-step cache entries, so each layer's `forward` (and
`ModelHelper.get_model_values`) finds its quantized weight, row range and log-range term ready.
AEWGS — the one estimator with an exchange step — is served too: one statistics launch, ONE packed
all-reduce for the whole model (the reference: three per weight tensor, gdnsq.py:126-129), one apply
launch.  Layers it cannot take (per-tensor, quantized bias, long rows, CPU) keep their own path; a
model driven without this call (e.g. the reference's own GDNSQQuant on these layer classes) simply
quantizes layer by layer.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .... import ops
from .. import _funnel
from .gdnsq_conv2d import NoisyConv2d


_DDP_GROUPS = 1


def prequantize_weights(model: torch.nn.Module) -> int:
    """Returns the number of layers served by the multi-tensor launch."""
    if not torch.is_grad_enabled():
        return 0                      # no_grad: the per-layer cache already survives across batches
    groups = {}
    for m in model.modules():
        if isinstance(m, NoisyConv2d) and m.training and m.multi_ok():
            key, hit = m.cache_probe()
            if hit is None:
                groups.setdefault((m.Q._method().value, m.weight.device), []).append((m, key))
    served = 0
    # One autograd node returns the gradients of ALL its weights when its LAST consumer's backward
    # has run, i.e. at the very end of the backward pass; under DDP the conv-weight buckets then
    # travel after the backward instead of under it.  Splitting the layers into a few consecutive
    # groups (deepest group ready first) would restore the overlap — measured at 2 GPUs with 4 groups:
    # ResNet-18 27.42 vs 27.43 ms / step, no effect (47 MB of gradients are ~0.1 ms of NVLink time) —
    # so ONE group, and for AEWGS ONE statistics all-reduce per step, is kept (_DDP_GROUPS).
    n_groups = _DDP_GROUPS if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else 1
    for (method, _dev), items in groups.items():
        if len(items) < 2:
            continue                  # a single layer gains nothing over its own launch
        per = -(-len(items) // n_groups)
        for g0 in range(0, len(items), per):
            part = items[g0:g0 + per]
            funnel = _funnel.active()
            res = ops.weight_fake_quant_rows_multi([m.weight for m, _ in part],
                                                   [m.log_wght_s for m, _ in part], method=method,
                                                   funnel=funnel)
            for (m, key), out in zip(part, res):
                m.adopt_quantized(key, *out)
        served += len(items)
    return served
