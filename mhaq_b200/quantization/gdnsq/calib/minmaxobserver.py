"""Calibration — mirror of the reference's src/quantization/gdnsq/calib/minmaxobserver.py:11-88.

`MinMaxObserver` records the running min / max of every NoisyAct input;
`apply_mean_stats_activations` turns them into (act_b, log_act_s, log_act_q);
`apply_quantile_weights_s` raises every weight scale to at least range / (2^bits - 1).
SURVEY.md §8 row (f)-2.  Differences that do not change results: the input's min / max come
out of the NoisyAct eval forward kernel itself (fused epilogue, `mhaq_fq_minmax_finalize`; one
`aminmax` pass as fall-back for a disabled quantizer) instead of separate `min` and `max` passes
plus a `torch.cat` per batch, and running min/max tensors instead of growing lists (the reference
only ever takes `.min()` / `.max()` of them).
Like the reference, calibration RE-BINDS the parameters to new `nn.Parameter` objects
(minmaxobserver.py:59-66, 86) — the layers' weight cache is keyed to survive that.
"""
import torch

from ..layers.gdnsq_act import NoisyAct
from ..layers.gdnsq_conv2d import NoisyConv2d
from ..layers.gdnsq_linear import NoisyLinear


class ObserverHook:
    def __call__(self, layer_name=None, *args):
        raise NotImplementedError("You need to implement __call__ method!")


class MinMaxObserver(ObserverHook):
    def __call__(self, module, input, output):
        return self._hook(module, input, output)

    def _hook(self, module, input, output) -> None:
        x = input[0]
        fused = getattr(module, "_in_minmax", None)
        if fused is not None and fused[:3] == (x.data_ptr(), x._version, tuple(x.shape)):
            # the eval forward kernel of this NoisyAct already reduced this very tensor
            # (min / max input ride along with min / max code): no extra pass
            mn, mx = fused[3][0:1], fused[3][1:2]
        else:
            mm = x.detach().aminmax()
            mn, mx = mm.min.reshape(1), mm.max.reshape(1)
        prev_mn = getattr(module, "min_values", None)
        if prev_mn is None or prev_mn.numel() == 0:
            module.min_values, module.max_values = mn, mx
        else:
            module.min_values = torch.minimum(prev_mn.to(mn.device), mn)
            module.max_values = torch.maximum(module.max_values.to(mx.device), mx)


def apply_mean_stats_activations(module, abits=8, max_bits=24):
    for _, m in module.named_modules():
        if not isinstance(m, NoisyAct):
            continue
        mn, mx = m.min_values.min(), m.max_values.max()
        m.min_values, m.max_values = torch.Tensor([]), torch.Tensor([])
        if not m.log_act_q.requires_grad and not m.log_act_s.requires_grad:
            abits = max_bits                       # (sticky, exactly like minmaxobserver.py:52-53)
        dev = m.log_act_s.device
        if mx - mn > 0:
            log_s = torch.log2((mx - mn) / (2 ** abits - 1))
            vals = ((m, "act_b", mn, m.act_b.requires_grad),
                    (m, "log_act_q", log_s + abits, m.log_act_q.requires_grad),
                    (m, "log_act_s", log_s, m.log_act_s.requires_grad))
        else:                                      # zero-width input: pruned, frozen
            # (reference quirk kept: `torch.tensor([0])` makes these two INTEGER parameters,
            # minmaxobserver.py:63-64 — exp2 of them is 1.0 either way)
            vals = ((m, "log_act_q", 0, False), (m, "log_act_s", 0, False), (m, "act_b", mn, False))
        for mod, name, v, rg in vals:
            t = torch.tensor([v], device=dev) if isinstance(v, int) else torch.tensor([float(v)], device=dev)
            setattr(mod, name, torch.nn.Parameter(t, requires_grad=rg))


def apply_quantile_weights_s(module, wbits=8, max_bits=24, qscheme="per-channel"):
    for _, m in module.named_modules():
        if not isinstance(m, (NoisyLinear, NoisyConv2d)):
            continue
        w = m.weight.detach()
        dims = tuple(range(1, w.dim()))
        max_, min_ = w.amax(dims), w.amin(dims)
        if not m.log_wght_s.requires_grad:
            wbits = max_bits
        log_s = torch.max(m.log_wght_s, torch.log2((max_ - min_) / (2 ** wbits - 1)).reshape(m.log_wght_s.shape))
        m.log_wght_s = torch.nn.Parameter(log_s.detach(), requires_grad=m.log_wght_s.requires_grad)
