"""CPU: the plugin surface (model surgery, state-dict layout, config plumbing) needs no GPU —
no kernel is launched until a forward pass."""
import pytest
import torch
from torch import nn

from mhaq_b200 import harness
from mhaq_b200.aux.types import QScheme
from mhaq_b200.quantization.quantizer import Quantizer
from mhaq_b200.quantization.gdnsq.gdnsq_utils import QNMethod
from mhaq_b200.quantization.gdnsq.layers.gdnsq_act import NoisyAct
from mhaq_b200.quantization.gdnsq.layers.gdnsq_conv2d import NoisyConv2d


def _quantized(model_name, **kw):
    lm = harness.LModule(harness.build_model(model_name), nn.CrossEntropyLoss())
    cfg = harness.make_config(excluded_layers=harness.EXCLUDED[model_name], **kw)
    return Quantizer(cfg)().quantize(lm, in_place=True)


def test_resnet20_surgery_and_state_dict_layout():
    q = _quantized("resnet20", qnmethod="AEWGS")
    convs = [(n, m) for n, m in q.model.named_modules() if isinstance(m, NoisyConv2d)]
    acts = [m for m in q.model.modules() if isinstance(m, NoisyAct)]
    assert len(convs) == 18 and len(acts) == 18
    assert isinstance(q.model.conv1, nn.Conv2d) and not isinstance(q.model.conv1, NoisyConv2d)
    assert isinstance(q.model.linear, nn.Linear)
    sd = q.model.state_dict()
    name = "layer1.0.conv1"
    for k, shape in ((f"{name}.activations_quantizer.log_act_q", (1,)),
                     (f"{name}.activations_quantizer.act_b", (1,)),
                     (f"{name}.activations_quantizer.log_act_s", (1,)),
                     (f"{name}.0.weight", (16, 16, 3, 3)),
                     (f"{name}.0.log_wght_s", (16, 1, 1, 1)),
                     (f"{name}.0.log_b_s", (1,)),
                     (f"{name}.0._noise_ratio", (1,))):
        assert tuple(sd[k].shape) == shape, k
    # init values (gdnsq_quant.py:532, gdnsq_act.py:12-25)
    assert float(sd[f"{name}.0.log_wght_s"].flatten()[0]) == -12
    assert float(sd[f"{name}.activations_quantizer.log_act_s"]) == -10
    assert float(sd[f"{name}.activations_quantizer.log_act_q"]) == 10
    # weights use the configured estimator, activations always STE (reference quirk 1)
    assert convs[0][1].Q.qnmethod == QNMethod.AEWGS and acts[0].Q.qnmethod == QNMethod.STE
    assert convs[0][1].qscheme == QScheme.PER_CHANNEL
    # the in-tree CIFAR ResNet uses functional relu: every activation is signed
    assert all(a.signed for a in acts)
    assert hasattr(q, "wrapped_criterion") and q.wrapped_criterion.at == 4 and q.wrapped_criterion.wt == 4


def test_resnet18_signedness_and_1x1_skipped():
    q = _quantized("resnet18", distillation=True)
    convs = {n: m for n, m in q.model.named_modules() if isinstance(m, NoisyConv2d)}
    assert len(convs) == 16                                # 3 downsample 1x1 convs skipped
    assert isinstance(q.model.layer2[0].downsample[0], nn.Conv2d)
    assert not isinstance(q.model.layer2[0].downsample[0], NoisyConv2d)
    # torchvision registers `relu` before conv2 in named_modules order: conv2 unsigned, conv1 signed
    b = q.model.layer1[0]
    assert b.conv1.activations_quantizer.signed and not b.conv2.activations_quantizer.signed
    assert b.conv2.activations_quantizer.act_b.requires_grad is False
    assert hasattr(q, "tmodel") and not any(p.requires_grad for p in q.tmodel.parameters())
    # weight / bias Parameters are shared with the original conv, not copied
    n_fp = sum(p.numel() for p in harness.build_model("resnet18").parameters())
    n_q = sum(p.numel() for n, p in q.model.named_parameters()
              if not any(t in n for t in ("log_", "act_b", "_noise_ratio")))
    assert n_fp == n_q


def test_excluded_layer_must_exist():
    lm = harness.LModule(harness.build_model("resnet20"), nn.CrossEntropyLoss())
    cfg = harness.make_config(excluded_layers=["nope"])
    with pytest.raises(AttributeError, match="not found"):
        Quantizer(cfg)().quantize(lm)


def test_unknown_estimator_name():
    lm = harness.LModule(harness.build_model("resnet20"), nn.CrossEntropyLoss())
    cfg = harness.make_config(excluded_layers=["conv1", "linear"], qnmethod="FOO")
    with pytest.raises(KeyError):
        Quantizer(cfg)().quantize(lm)


def test_rounding_noise_functions_route_to_the_kernels_and_refuse_cpu_tensors():
    # QNSTE & co. (gdnsq.py:32-147) stay callable: they run the kernels (ops.rounding_noise);
    # like every other entry point they refuse CPU tensors instead of falling back
    from mhaq_b200.quantization.gdnsq.gdnsq import QNSTE, QNLSQ, QNEWGS, QNAEWGS, scaled_noise
    for fn in (QNSTE.apply, QNLSQ.apply, QNEWGS.apply, QNAEWGS.apply, scaled_noise):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn(torch.zeros(2), torch.ones(1))


def test_layers_accept_the_reference_packages_own_enums():
    # the reference's GDNSQQuant constructs these classes with ITS enums
    import enum
    RefQScheme = enum.Enum("QScheme", {"PER_TENSOR": 0, "PER_CHANNEL": 1})
    RefQN = enum.Enum("QNMethod", {"STE": 0, "EWGS": 1, "AEWGS": 2, "LSQ": 3})
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_conv2d import NoisyConv2d
    c = NoisyConv2d(4, 8, 3, qscheme=RefQScheme.PER_CHANNEL, qnmethod=RefQN.LSQ)
    assert c.log_wght_s.shape == (8, 1, 1, 1) and hasattr(c, "log_b_s")
    assert c.qscheme is RefQScheme.PER_CHANNEL                # stored as given
    assert c.Q._method().name == "LSQ"
    c = NoisyConv2d(4, 8, 3, qscheme=RefQScheme.PER_TENSOR, qnmethod=RefQN.STE)
    assert c.log_wght_s.shape == (1,)
    Bad = enum.Enum("QNMethod", {"STE": 7})
    c.Q.qnmethod = Bad.STE
    with pytest.raises(AttributeError, match="Unknown method"):
        c.Q._method()


def test_src_alias_resolves_the_reference_import_paths():
    # run in a fresh interpreter: the alias must not collide with a loaded reference `src`
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import mhaq_b200.compat as c; names = c.install_src_alias()\n"
        "from src.quantization.quantizer import Quantizer\n"
        "from src.quantization.gdnsq.gdnsq_quant import GDNSQQuant\n"
        "from src.quantization.gdnsq.layers.gdnsq_act import NoisyAct\n"
        "from src.quantization.gdnsq.layers.gdnsq_conv2d import NoisyConv2d\n"
        "from src.quantization.gdnsq.layers.gdnsq_linear import NoisyLinear\n"
        "from src.quantization.gdnsq.utils.model_helper import ModelHelper\n"
        "from src.quantization.gdnsq.utils import model_stats\n"
        "from src.quantization.gdnsq.calib.minmaxobserver import MinMaxObserver, apply_mean_stats_activations, apply_quantile_weights_s\n"
        "from src.quantization.gdnsq.calib.hooks import register_lightning_activation_forward_hook\n"
        "from src.quantization.gdnsq.config.config_schema import GDNSQQuantizerParams\n"
        "from src.quantization.gdnsq.gdnsq import Quantizer as Q, QNoise, QNSTE, QNLSQ, QNEWGS, QNAEWGS, reduce_to_shape, scaled_noise\n"
        "from src.quantization.gdnsq.gdnsq_utils import QNMethod, QMode\n"
        "from src.aux.types import QScheme, QMethod, MType, DType\n"
        "from src.aux.qutils import attrsetter, is_biased\n"
        "import src.quantization as pkg, mhaq_b200.quantization as real\n"
        "assert pkg is real and pkg.GDNSQQuant is GDNSQQuant\n"
        "import mhaq_b200.quantization.gdnsq.layers.gdnsq_act as a; assert a.NoisyAct is NoisyAct\n"
        "print('ok', len(names))\n" % root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stderr[-1500:]


def test_cpu_forward_fails_loudly():
    act = NoisyAct()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        act(torch.randn(4, 4))


def test_calibration_api_rebinds_parameters_like_the_reference():
    from mhaq_b200.quantization.gdnsq.calib.minmaxobserver import (MinMaxObserver, apply_mean_stats_activations,
                                                                   apply_quantile_weights_s)
    q = _quantized("resnet20")
    conv = q.model.layer1[0].conv1[1] if False else q.model.layer1[0].conv1._modules["0"]
    act = q.model.layer1[0].conv1.activations_quantizer
    old_ws, old_as = conv.log_wght_s, act.log_act_s
    apply_quantile_weights_s(q.model, wbits=10)
    assert conv.log_wght_s is not old_ws and conv.log_wght_s.requires_grad
    w = conv.weight.detach()
    expect = torch.log2((w.amax((1, 2, 3)) - w.amin((1, 2, 3))) / (2 ** 10 - 1)).reshape(-1, 1, 1, 1)
    assert torch.allclose(conv.log_wght_s.detach(), torch.max(torch.full_like(expect, -12.0), expect))
    obs = MinMaxObserver()
    x1, x2 = torch.randn(2, 16, 8, 8), torch.randn(2, 16, 8, 8) * 3
    for m in q.model.modules():
        if isinstance(m, NoisyAct):
            obs(m, (x1,), None)
            obs(m, (x2,), None)
    apply_mean_stats_activations(q.model, abits=10)
    mn, mx = torch.minimum(x1.min(), x2.min()), torch.maximum(x1.max(), x2.max())
    assert act.log_act_s is not old_as
    assert torch.allclose(act.act_b.detach(), mn.reshape(1))
    log_s = torch.log2((mx - mn) / (2 ** 10 - 1))
    assert torch.allclose(act.log_act_s.detach(), log_s.reshape(1))
    assert torch.allclose(act.log_act_q.detach(), (log_s + 10).reshape(1))
    assert act.act_b.requires_grad and act.log_act_s.requires_grad


def test_rfdn_surgery_matches_the_reference_inventory():
    """BASELINE configs[4]: RFDN(scale=4) with excluded
    ['fea_conv', 'upsampler.0'] -> 33 quantized 3x3 convolutions (358 236 weight elements),
    the 29 1x1 convolutions untouched, every activation quantizer signed, L1 criterion wrapped
    by PotentialLossNoPred (no teacher)."""
    from mhaq_b200 import harness
    from mhaq_b200.quantization.gdnsq.gdnsq_loss import PotentialLossNoPred
    from mhaq_b200.quantization.gdnsq.gdnsq_utils import QNMethod
    q = harness.build_qat("rfdn", "cpu", qnmethod="LSQ", act_bit=2, weight_bit=2, distillation=False)
    convs = [m for m in q.model.modules() if isinstance(m, NoisyConv2d)]
    assert len(convs) == 33 and sum(c.weight.numel() for c in convs) == 358236
    assert all(c.kernel_size == (3, 3) and c.Q.qnmethod == QNMethod.LSQ for c in convs)
    plain = [m for m in q.model.modules() if type(m) is nn.Conv2d]
    assert len(plain) == 29 + 2 and sum(m.kernel_size == (1, 1) for m in plain) == 29
    assert type(q.model.fea_conv) is nn.Conv2d and type(q.model.upsampler[0]) is nn.Conv2d
    acts = [m for m in q.model.modules() if isinstance(m, NoisyAct)]
    assert len(acts) == 33 and all(a.signed for a in acts)
    assert isinstance(q.wrapped_criterion, PotentialLossNoPred) and not hasattr(q, "tmodel")
    assert sum(p.numel() for p in harness.Rfdn().parameters()) == 433448
