"""Gradient funnel: ONE gradient per quantizer parameter per step.

`log_act_s`, `log_act_q` and `log_wght_s` are read twice in a QAT step: by the fake-quant kernels and
by `PotentialLoss` (through `ModelHelper.get_model_values`, utils/model_helper.py:13-76).  Autograd
therefore accumulates two gradients per parameter — a tiny `add` kernel each, 54 launches per
ResNet-20 step.  While a funnel step is active the quantizer ops return ALIASES of those parameters
as extra outputs of their own autograd node; `ModelHelper` hands the aliases (not the leaves) to the
loss, so the loss's gradient flows back into the quantizer node, whose backward kernel adds it to
the gradient it emits (`mhaq_fq_bwd_fused_acc_f32`, `g_log_scale_acc`): the same fp32 sum autograd
would have formed, without the launches.

Only this package's `GDNSQQuant` training steps activate it; with the reference's own plugin on
top of these layers nothing changes.
"""
import contextlib

_active = False
_holders = []


def active() -> bool:
    return _active


def hold(module) -> None:
    """`module` stored aliases of this step (`_funnel_act` / `_funnel_ls`); they are dropped when
    the step ends, consumed or not, so no module keeps a finished step's graph alive."""
    _holders.append(module)


@contextlib.contextmanager
def step():
    global _active
    prev, _active = _active, True
    try:
        yield
    finally:
        _active = prev
        if not prev:
            for m in _holders:
                m.__dict__.pop("_funnel_act", None)
                m.__dict__.pop("_funnel_ls", None)
            _holders.clear()
