"""Drop-in proof on the GPU: the reference's OWN upper layers — `Quantizer(config)().quantize`,
`GDNSQQuant`'s patched training step, `ModelHelper.get_model_values`, `PotentialLoss`, the
calibration functions, `model_stats` — all imported unmodified from oracle/_ref, run on top of this
repo's layer classes (swapped in per INTEGRATION.md §B) and are compared, step for step, with the
same pipeline on the reference's own layers on the same GPU.
"""
import copy
import math
from collections import OrderedDict

import pytest
import torch
from torch import nn

from oracle import checks as C
from oracle import ref_harness as RH
from oracle import ref_loader

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not ref_loader.available():
        pytest.skip("oracle/_ref not staged")
    return ref_loader.load_full()


class _Net(nn.Module):
    """Every quantized activation / weight has a distinct shape, so the recorded noise of a
    quantizer can be matched by shape in both pipelines."""

    def __init__(self, n_cls=10):
        super().__init__()
        self.stem = nn.Conv2d(3, 8, 3, padding=1, bias=False)           # excluded (like conv1)
        self.bn0 = nn.BatchNorm2d(8)
        self.relu0 = nn.ReLU()
        self.c1 = nn.Conv2d(8, 12, 3, padding=1, bias=True)             # after nn.ReLU: unsigned
        self.bn1 = nn.BatchNorm2d(12)
        self.c2 = nn.Conv2d(12, 16, 3, stride=2, padding=1, bias=False)  # after BatchNorm: signed
        self.relu2 = nn.ReLU()
        self.c3 = nn.Conv2d(16, 20, 3, stride=2, padding=1, bias=False)
        self.skip = nn.Conv2d(20, 20, 1, bias=False)                    # 1x1: never quantized
        self.head = nn.Linear(20, n_cls)                                # excluded

    def forward(self, x):
        x = self.relu0(self.bn0(self.stem(x)))
        x = self.bn1(self.c1(x))
        x = self.relu2(self.c2(x))
        x = self.skip(self.c3(x))
        return self.head(x.mean((2, 3)))


class _ShapeNoise:
    """One fixed {-0.5,+0.5} draw per tensor shape, served to the reference through
    torch.randint_like and to the product through the kernels' explicit-noise argument."""

    def __init__(self, device, seed=0):
        self.gen = torch.Generator(device="cpu").manual_seed(seed)
        self.device, self.table = device, {}
        self._orig = torch.randint_like

    def get(self, shape):
        shape = tuple(shape)
        if shape not in self.table:
            self.table[shape] = (torch.randint(0, 2, shape, generator=self.gen).float() - 0.5).to(self.device)
        return self.table[shape]

    def __enter__(self):
        def fake(t, high, **kw):
            return (self.get(t.shape) + 0.5).to(t.dtype)
        torch.randint_like = fake
        return self

    def __exit__(self, *a):
        torch.randint_like = self._orig


@pytest.mark.parametrize("method", ["STE", "LSQ", "AEWGS"])
def test_reference_pipeline_on_swapped_layers_matches_its_own_layers(ref, method):
    from mhaq_b200 import ops
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_act import NoisyAct
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_conv2d import NoisyConv2d
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_linear import NoisyLinear
    dev = torch.device("cuda")
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    base = _Net()
    x = torch.randn(32, 3, 24, 24, device=dev)
    t = torch.randint(0, 10, (32,), device=dev)
    cfg = RH.make_cfg(ref, act_bit=4, weight_bit=4, qscheme=1, qnmethod=method,
                      excluded_layers=["stem", "head"], distillation=True,
                      distillation_loss="Symmetrical KL")
    noise = _ShapeNoise(dev)

    def build(swapped):
        lm = RH.build_lmodule(ref, copy.deepcopy(base), 10, lr=1e-3).to(dev)
        q = RH.quantize(ref, lm, cfg).to(dev)
        RH.calibrate(ref, q, x, act_bits=4, weight_bits=4, device=dev)
        return q

    # --- the reference on its own layers
    q_ref = build(False)
    grads_ref = {}
    with noise:
        losses, opt_ref = RH.train_steps(q_ref, (x, t), 1, on_step=lambda i, m: grads_ref.update(
            {n: p.grad.clone() for n, p in m.model.named_parameters() if p.grad is not None}))
    loss_ref = losses[0]

    # --- the same reference code on the product's layers
    real_act, real_rows, real_fq = ops.act_fake_quant, ops.weight_fake_quant_rows, ops.fake_quant

    def act_with_noise(x_, ls, lq, b, method="STE", noise=None, philox=None):
        return real_act(x_, ls, lq, b, method=method, noise=noise_tab(x_.shape))

    def rows_with_noise(w, ls, method="STE", noise=None, philox=None):
        return real_rows(w, ls, method=method, noise=None if ops._method_id(method) == 3 else noise_tab(w.shape))

    def fq_with_noise(*a, **k):
        if ops._method_id(k.get("method", "STE")) != 3:
            k["noise"] = noise_tab(a[0].shape)
        return real_fq(*a, **k)

    def wlog_with_noise(w, ls, method="STE", noise=None, philox=None):
        return real_wlog(w, ls, method=method, noise=None if ops._method_id(method) == 3 else noise_tab(w.shape))

    real_wlog = ops.weight_fake_quant_log
    noise_tab = noise.get
    grads_new = {}
    with RH.swapped_layers(ref, NoisyAct, NoisyConv2d, NoisyLinear):
        q_new = build(True)
        assert any(isinstance(m, NoisyConv2d) for m in q_new.model.modules())
        # calibration (the reference's functions, forward hooks on the product's NoisyAct) agrees
        for (n1, p1), (n2, p2) in zip(q_ref.model.named_parameters(), q_new.model.named_parameters()):
            assert n1 == n2
        ops.act_fake_quant, ops.weight_fake_quant_rows = act_with_noise, rows_with_noise
        ops.fake_quant, ops.weight_fake_quant_log = fq_with_noise, wlog_with_noise
        try:
            losses, opt_new = RH.train_steps(q_new, (x, t), 1, on_step=lambda i, m: grads_new.update(
                {n: p.grad.clone() for n, p in m.model.named_parameters() if p.grad is not None}))
        finally:
            ops.act_fake_quant, ops.weight_fake_quant_rows = real_act, real_rows
            ops.fake_quant, ops.weight_fake_quant_log = real_fq, real_wlog
        loss_new = losses[0]
        # validation-time statistics of the reference, on the product's layers
        q_new.eval(); q_ref.eval()
        with torch.no_grad():
            out_new = q_new.model(x)
        w_new = ref.model_stats.get_true_weights_width(q_new.model)
        a_new = ref.model_stats.get_true_activations_width(q_new.model)
    with torch.no_grad():
        out_ref = q_ref.model(x)
    w_ref = ref.model_stats.get_true_weights_width(q_ref.model)
    a_ref = ref.model_stats.get_true_activations_width(q_ref.model)
    torch.backends.cudnn.allow_tf32 = tf32

    # forward is bit-exact layer by layer, so the loss agrees to rounding of its own reductions;
    # gradients: STE / LSQ input gradients are bit-exact (only the parameter-gradient sums differ in
    # summation order), AEWGS input gradients agree to 1e-5 per layer and compound through the net
    C.assert_close_rel(loss_new, loss_ref, 1e-6, "loss")
    assert set(grads_new) == set(grads_ref)
    rel = 2e-4 if method == "AEWGS" else 2e-5
    for n in sorted(grads_ref):
        g_r, g_n = grads_ref[n], grads_new[n]
        scale = float(g_r.abs().max())
        C.assert_close_rel(g_n, g_r, rel, f"grad {n}", abs_floor=rel * scale + 1e-8)
    # after the optimizer step both models still agree (the update consumed matching gradients)
    C.assert_close_rel(out_new, out_ref, 1e-3, "eval output after one step", abs_floor=1e-4)
    C.assert_close_rel(torch.as_tensor(float(w_new)), torch.as_tensor(float(w_ref)), 1e-4, "true weight width", abs_floor=1e-4)
    C.assert_close_rel(torch.as_tensor(float(a_new)), torch.as_tensor(float(a_ref)), 1e-4, "true activation width", abs_floor=1e-4)


def test_product_plugin_through_the_src_alias_trains_on_gpu():
    """`from src.quantization.quantizer import Quantizer` (the reference's import path, resolved by
    mhaq_b200.compat) drives a QAT step on the GPU — in a process of its own, since this test
    session also holds the live reference's `src` package."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys; sys.path.insert(0, {root!r})
import torch
import mhaq_b200.compat as compat
compat.install_src_alias()
from src.quantization.quantizer import Quantizer
from src.quantization.gdnsq.utils import model_stats
from src.aux.types import QScheme
from mhaq_b200 import harness
dev = torch.device('cuda')
lm = harness.LModule(harness.build_model('resnet20', 10).to(dev), torch.nn.CrossEntropyLoss(), torch.optim.RAdam, 1e-3)
cfg = harness.make_config(act_bit=4, weight_bit=4, qscheme=1, qnmethod='STE', excluded_layers=('conv1', 'linear'))
q = Quantizer(cfg)().quantize(lm, in_place=True).to(dev)
x = torch.randn(64, 3, 32, 32, device=dev); t = torch.randint(0, 10, (64,), device=dev)
harness.calibrate(q, x)
opt = q.configure_optimizers(); q.train(); q.wrapped_criterion.train()
l0 = None
for i in range(3):
    loss = q.training_step((x, t), i); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
    l0 = l0 or float(loss)
assert torch.isfinite(loss)
print('ok', l0, float(loss), float(model_stats.get_true_weights_width(q.model)))
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_calibration_and_model_stats_mirrors_match_the_reference(ref):
    """Rows (f)-2 / (f)-3: this repo's mirrors of calib/minmaxobserver.py:19-88 (with the input
    min / max coming out of the eval forward kernel) and utils/model_stats.py:116-262 (tensor ops
    instead of the per-channel Python loop) against the reference's own functions on the
    reference's own layers — same model, same calibration batches, same GPU."""
    from mhaq_b200 import harness
    from mhaq_b200.quantization.quantizer import Quantizer as OurFactory
    from mhaq_b200.quantization.gdnsq.calib import minmaxobserver as our_mm
    from mhaq_b200.quantization.gdnsq.calib.hooks import register_lightning_activation_forward_hook as our_hook
    from mhaq_b200.quantization.gdnsq.utils import model_stats as our_stats
    dev = torch.device("cuda")
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(1)
        base = _Net()
        with torch.no_grad():                       # a dead quantized input -> the "pruned" branch (:62-66)
            base.c2.weight.zero_(); base.c2.weight[:, :, 1, 1] = 0.0
            base.bn1.weight.zero_(); base.bn1.bias.zero_()
        batches = [torch.randn(16, 3, 24, 24, device=dev) * (1 + i) for i in range(3)]
        cfg = RH.make_cfg(ref, act_bit=4, weight_bit=4, qscheme=1, qnmethod="STE",
                          excluded_layers=["stem", "head"])
        q_ref = RH.quantize(ref, RH.build_lmodule(ref, copy.deepcopy(base), 10).to(dev), cfg).to(dev)
        q_our = OurFactory(cfg)().quantize(
            harness.LModule(copy.deepcopy(base).to(dev), nn.CrossEntropyLoss(), torch.optim.RAdam, 1e-3), in_place=True).to(dev)

        # --- weights: apply_quantile_weights_s (:69-88)
        ref.minmaxobserver.apply_quantile_weights_s(q_ref.model, wbits=5)
        our_mm.apply_quantile_weights_s(q_our.model, wbits=5)
        # --- activations: MinMaxObserver hooks over several batches, then apply_mean_stats_activations (:39-66)
        obs = ref.minmaxobserver.MinMaxObserver.__new__(ref.minmaxobserver.MinMaxObserver)
        h_ref = ref.hooks.register_lightning_activation_forward_hook(q_ref.model, obs)
        h_our = our_hook(q_our.model, our_mm.MinMaxObserver())
        q_ref.eval(); q_our.eval()
        with torch.no_grad():
            for b in batches:
                q_ref.model(b); q_our.model(b)
        for h in h_ref + h_our:
            h.remove()
        # the fused epilogue really served the hook (no extra reduction pass on our side)
        from mhaq_b200.quantization.gdnsq.layers.gdnsq_act import NoisyAct
        assert all(m._in_minmax is not None for m in q_our.model.modules() if isinstance(m, NoisyAct))
        ref.minmaxobserver.apply_mean_stats_activations(q_ref.model, abits=6)
        our_mm.apply_mean_stats_activations(q_our.model, abits=6)
        q_ref.to(dev); q_our.to(dev)
        p_ref, p_our = dict(q_ref.model.named_parameters()), dict(q_our.model.named_parameters())
        assert list(p_ref) == list(p_our)
        n_frozen = 0
        for n in p_ref:
            assert p_ref[n].requires_grad == p_our[n].requires_grad, n
            assert torch.equal(p_ref[n].detach().cpu(), p_our[n].detach().cpu()), n
            n_frozen += int(n.endswith("log_act_s") and not p_ref[n].requires_grad)
        assert n_frozen >= 1                        # the pruned branch was exercised

        # --- validation-time statistics
        with torch.no_grad():
            q_ref.model(batches[0]); q_our.model(batches[0])     # sets NoisyAct.bw (gdnsq_act.py:51-54)
        S = ref.model_stats
        pairs = [
            (S.get_true_weights_width(q_ref.model), our_stats.get_true_weights_width(q_our.model)),
            (S.get_true_weights_width(q_ref.model, max=False), our_stats.get_true_weights_width(q_our.model, max=False)),
            (S.get_true_activations_width(q_ref.model), our_stats.get_true_activations_width(q_our.model)),
            (S.get_true_activations_width(q_ref.model, max=False), our_stats.get_true_activations_width(q_our.model, max=False)),
            (S.get_weights_bit_width_mean(q_ref.model), our_stats.get_weights_bit_width_mean(q_our.model)),
        ]
        # a pruned activation's log parameters are INTEGER tensors in the reference
        # (minmaxobserver.py:63-64), on which its own `.mean()` raises: same parameters, same error here
        for name in ("log_act_s", "log_act_q"):
            dt_r = {n: p.dtype for n, p in p_ref.items() if n.endswith(name)}
            dt_o = {n: p.dtype for n, p in p_our.items() if n.endswith(name)}
            assert dt_r == dt_o and torch.int64 in dt_r.values()
        with pytest.raises(RuntimeError):
            S.get_activations_bit_width_mean(q_ref.model)
        with pytest.raises(RuntimeError):
            our_stats.get_activations_bit_width_mean(q_our.model)
        for i, (a, b) in enumerate(pairs):
            assert math.isclose(float(a), float(b), rel_tol=1e-6, abs_tol=1e-6), (i, float(a), float(b))
        for m_r, m_o in zip(q_ref.model.modules(), q_our.model.modules()):
            if isinstance(m_r, ref.NoisyConv2d):
                for mx in (True, False):
                    assert math.isclose(float(S.get_true_layer_bit_width(m_r, max=mx)),
                                        float(our_stats.get_true_layer_bit_width(m_o, max=mx)), rel_tol=1e-6, abs_tol=1e-6)
        q_ref.wrapped_criterion.wt = q_our.wrapped_criterion.wt = 8
        q_ref.wrapped_criterion.at = q_our.wrapped_criterion.at = 8
        assert bool(S.is_converged(q_ref)) == bool(our_stats.is_converged(q_our))
        q_ref.wrapped_criterion.wt = q_our.wrapped_criterion.wt = 1
        assert bool(S.is_converged(q_ref)) == bool(our_stats.is_converged(q_our))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
