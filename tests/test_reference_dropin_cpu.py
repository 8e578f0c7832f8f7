"""Drop-in proof, CPU part (no kernels run: model surgery, configs, state-dict contract).

The LIVE reference (oracle/_ref) is imported next to the product and both are driven with the
`quantization:` section of the three YAML configs BASELINE.json names:
  * this repo's plugin package (`mhaq_b200.quantization`, the `src.quantization` alias of
    mhaq_b200.compat) and the reference's own `Quantizer(config)().quantize(lmodel)` must produce
    the same module tree, the same state-dict keys / shapes / initial values and the same
    signedness / estimator / bit-width plumbing;
  * the reference's OWN `GDNSQQuant.quantize`, with this repo's layer classes swapped in per
    INTEGRATION.md §B, must build a model made of the product's layers with that same layout.
"""
import copy

import pytest
import torch
from torch import nn

from oracle import ref_harness as RH
from oracle import ref_loader


@pytest.fixture(scope="module")
def ref():
    if not ref_loader.available():
        pytest.skip("oracle/_ref not staged")
    return ref_loader.load_full()


YAMLS = {
    "resnet18": "gdnsq_config_resnet18_imagenet_ste_w4a4.yaml",
    "resnet20": "gdnsq_config_resnet20_cifar100_aewgs_w1a1.yaml",
    "rfdn": "gdnsq_config_rfdn_lsq_w2a2.yaml",
}


def _model(ref, name):
    torch.manual_seed(0)
    if name == "resnet18":
        import torchvision
        return torchvision.models.resnet18(num_classes=1000), 1000, None
    if name == "resnet20":      # pytorchcv (the YAML's model provider) is not in the image: the in-tree
        #                         CIFAR ResNet-20 with ITS first / last layer names excluded instead
        return ref.resnet_cifar.resnet20_cifar10(num_classes=100), 100, ["conv1", "linear"]
    from src.models.sr.rfdn.rfdn import RFDN
    return RFDN(), 10, None


def _product_lmodule(model, lr=3e-4):
    from mhaq_b200 import harness
    return harness.LModule(model, nn.CrossEntropyLoss(), torch.optim.RAdam, lr)


@pytest.mark.parametrize("name", list(YAMLS))
def test_yaml_configs_drive_both_plugins_to_the_same_model(ref, name):
    from mhaq_b200.quantization.quantizer import Quantizer as OurFactory
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_act import NoisyAct as OurAct
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_conv2d import NoisyConv2d as OurConv
    model, n_cls, excluded = _model(ref, name)
    cfg = RH.make_cfg(ref, YAMLS[name])
    if excluded is not None:
        cfg.quantization.excluded_layers = excluded
    q = cfg.quantization
    assert q.name == "GDNSQQuant"
    expect = {"resnet18": ("STE", 4, 4, True), "resnet20": ("AEWGS", 1, 1, True), "rfdn": ("LSQ", 2, 2, False)}[name]
    assert (q.params.qnmethod, q.act_bit, q.weight_bit, bool(q.params.distillation)) == expect

    theirs = RH.quantize(ref, RH.build_lmodule(ref, copy.deepcopy(model), n_cls), cfg)
    ours = OurFactory(cfg)().quantize(_product_lmodule(copy.deepcopy(model)), in_place=True)

    sd_t, sd_o = theirs.model.state_dict(), ours.model.state_dict()
    assert list(sd_t.keys()) == list(sd_o.keys())
    for k in sd_t:
        assert sd_t[k].shape == sd_o[k].shape and torch.equal(sd_t[k], sd_o[k]), k
    # same surgery: same modules replaced, same signedness, same estimator on the weights,
    # activations always STE (reference quirk 1)
    mods_t = dict(theirs.model.named_modules())
    n_conv = 0
    for n, m in ours.model.named_modules():
        if isinstance(m, OurConv):
            t = mods_t[n]
            assert isinstance(t, ref.NoisyConv2d)
            assert m.Q._method().name == t.Q.qnmethod.name == q.params.qnmethod
            assert m.qscheme.value == t.qscheme.value == 1
            n_conv += 1
        elif isinstance(m, OurAct):
            t = mods_t[n]
            assert isinstance(t, ref.NoisyAct)
            assert m.signed == t.signed and m.disable == t.disable
            assert m.Q._method().name == t.Q.qnmethod.name == "STE"
    assert n_conv == {"resnet18": 16, "resnet20": 18, "rfdn": 33}[name]
    assert (ours.wrapped_criterion.at, ours.wrapped_criterion.wt) == (theirs.wrapped_criterion.at, theirs.wrapped_criterion.wt)
    assert type(ours.wrapped_criterion).__name__ == type(theirs.wrapped_criterion).__name__
    # the reference's state dict loads strictly into the product's model and back
    ours.model.load_state_dict(sd_t, strict=True)
    theirs.model.load_state_dict(sd_o, strict=True)


@pytest.mark.parametrize("name", ["resnet18", "rfdn"])
def test_reference_gdnsqquant_builds_the_products_layers_when_swapped(ref, name):
    """INTEGRATION.md §B executed in memory: the reference's own GDNSQQuant.quantize (unmodified
    code, its own enums and config plumbing) constructs this repo's layer classes."""
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_act import NoisyAct
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_conv2d import NoisyConv2d
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_linear import NoisyLinear
    model, n_cls, excluded = _model(ref, name)
    cfg = RH.make_cfg(ref, YAMLS[name])
    plain = RH.quantize(ref, RH.build_lmodule(ref, copy.deepcopy(model), n_cls), cfg)
    with RH.swapped_layers(ref, NoisyAct, NoisyConv2d, NoisyLinear):
        swapped = RH.quantize(ref, RH.build_lmodule(ref, copy.deepcopy(model), n_cls), cfg)
        kinds = {type(m) for m in swapped.model.modules()}
        assert NoisyConv2d in kinds and NoisyAct in kinds
        assert ref.NoisyConv2d not in kinds and ref.NoisyAct not in kinds
        sd_p, sd_s = plain.model.state_dict(), swapped.model.state_dict()
        assert list(sd_p.keys()) == list(sd_s.keys())
        assert all(torch.equal(sd_p[k], sd_s[k]) for k in sd_p)
        # the reference's per-step collector walks the product's layers (isinstance on the swapped
        # names) — parameters only, no kernels: usable on the CPU
        vals = ref.ModelHelper.get_model_values(swapped.model, swapped.qscheme)
        # calibration of the weights (no forward pass): the reference's function on the product's layers
        ref.minmaxobserver.apply_quantile_weights_s(swapped.model, wbits=4)
    vals_p = ref.ModelHelper.get_model_values(plain.model, plain.qscheme)
    assert len(vals) == len(vals_p) == 4
    for a, b in zip(vals, vals_p):
        assert torch.equal(a, b)
    ref.minmaxobserver.apply_quantile_weights_s(plain.model, wbits=4)
    for (n1, p1), (n2, p2) in zip(plain.model.named_parameters(), swapped.model.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2), n1


def test_all_distillation_losses_of_the_reference_are_available_and_agree(ref):
    """gdnsq_quant.py:40-66: the eight `distillation_loss` names, values against the reference's."""
    from mhaq_b200.quantization.gdnsq.distill_losses import get_distillation_loss
    import types
    torch.manual_seed(0)
    s, t = torch.randn(16, 10), torch.randn(16, 10)
    for name in ("Cross-Entropy", "Symmetrical Cross-Entropy", "L1", "L2", "KL", "Hellinger", "Symmetrical KL", "JSD"):
        cfg = types.SimpleNamespace(quantization=types.SimpleNamespace(
            params=types.SimpleNamespace(distillation=True, distillation_loss=name)))
        theirs = ref.GDNSQQuant.get_loss(types.SimpleNamespace(config=cfg), qmodel=None)
        a, b = get_distillation_loss(name)(s, t), theirs(s, t)
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), (name, float(a), float(b))
