"""Import the LIVE reference (aifoundry-org/MHAQ, unmodified) from ``oracle/_ref/``.

TEST / BENCH INFRASTRUCTURE ONLY — nothing under ``mhaq_b200/`` imports it; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s reference / cpu_baseline legs may.

``oracle/make_ref.py`` stages the reference's ``src/`` tree and ``config/*.yaml`` under
``oracle/_ref/`` (git-ignored, shipped to the GPU box).  The reference imports itself as
``src.*`` and depends on packages this image does not have (Lightning, torchmetrics,
pytorchcv, piq, matplotlib, seaborn, wandb, DALI, pycocotools — SURVEY.md §8c).  This loader

* puts ``oracle/_ref`` on ``sys.path`` so ``import src...`` resolves to the staged reference,
* installs *stub* modules for the absent third-party packages (only for those that really are
  absent): ``lightning.pytorch.LightningModule`` is an ``nn.Module`` with a no-op ``log``,
  everything else is an inert placeholder class.  None of the stubs carries arithmetic: every
  number the reference produces comes from its own code and torch.

Two levels:
    ref = load_ops()   # ops + layer wrappers only (bypasses src/quantization/__init__.py)
    ref = load_full()  # the whole package: GDNSQQuant, ModelHelper, PotentialLoss, model_stats,
                       # calibration, in-tree models, LVisionCls ...
Both return a ``types.SimpleNamespace`` of the reference's modules / classes.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

# third-party packages the reference imports that this image lacks (probed)
_ABSENT_OK = ("lightning", "pytorch_lightning", "torchmetrics", "pytorchcv", "piq", "matplotlib",
              "seaborn", "wandb", "nvidia", "pycocotools", "tensorboard", "PIL", "cv2", "thop",
              "albumentations", "torchinfo")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "src", "quantization", "gdnsq", "gdnsq.py"))


def ref_dir() -> str:
    if not available():
        raise FileNotFoundError(
            f"{REF_DIR} is empty: run `python oracle/make_ref.py` in the build container "
            "(needs /root/reference); the staged copy then travels to the GPU box")
    return REF_DIR


# ---------------------------------------------------------------------------------------------
# stubs for absent third-party packages
# ---------------------------------------------------------------------------------------------
class _Inert:
    """Placeholder class: constructible, callable, subclassable, attribute access yields more."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()


class _LightningModule(nn.Module):
    """What GDNSQQuant / LVisionCls need from pl.LightningModule: an nn.Module with log sinks."""

    def __init__(self, *a, **k):
        super().__init__()
        self.logged = {}
        self.trainer = types.SimpleNamespace(logged_metrics={}, global_step=0, current_epoch=0)

    def log(self, name, value, *a, **k):
        self.logged[name] = value.detach() if torch.is_tensor(value) else value

    def log_dict(self, d, *a, **k):
        for n, v in dict(d).items():
            self.log(n, v)

    def save_hyperparameters(self, *a, **k):
        pass

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name == "LightningModule":
            return _LightningModule
        if name in ("rank_zero_only", "rank_zero_warn", "rank_zero_info"):
            return (lambda f=None, *a, **k: f) if name == "rank_zero_only" else (lambda *a, **k: None)
        if name == "_PATH":
            return str
        if name[0].islower():            # `from lightning.pytorch import loggers` -> a sub-module
            return importlib.import_module(f"{self.__name__}.{name}")
        cls = type(name, (_Inert,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        m.__mhaq_stub__ = True
        return m

    def exec_module(self, module):
        if module.__name__ == "matplotlib":
            module.scale = None          # the stray `from matplotlib import scale`, gdnsq.py:1


_finder = None


def install_stubs():
    """Stub ONLY what cannot be imported for real."""
    global _finder
    if _finder is not None:
        return sorted(_finder.roots)
    missing = []
    for name in _ABSENT_OK:
        if name in sys.modules and not getattr(sys.modules[name], "__mhaq_stub__", False):
            continue
        try:
            if importlib.util.find_spec(name) is None:
                missing.append(name)
        except (ImportError, ValueError):
            missing.append(name)
    _finder = _StubFinder(missing)
    sys.meta_path.append(_finder)
    return sorted(missing)


# ---------------------------------------------------------------------------------------------
def _check_src_is_reference():
    m = sys.modules.get("src")
    if m is None:
        return
    paths = [os.path.realpath(p) for p in getattr(m, "__path__", [])]
    if not any(p.startswith(os.path.realpath(REF_DIR)) for p in paths):
        raise RuntimeError("a different `src` package is already imported in this process "
                           f"({paths}); the live reference must be loaded in a process of its own")


def load_ops():
    """Ops + layer wrappers of the live reference, without src/quantization/__init__.py (which
    pulls the whole model zoo): SURVEY.md Appendix D's recipe against the staged tree."""
    root = ref_dir()
    install_stubs()
    _check_src_is_reference()
    for name, sub in [("src", "src"), ("src.quantization", "src/quantization"),
                      ("src.quantization.gdnsq", "src/quantization/gdnsq"),
                      ("src.quantization.gdnsq.layers", "src/quantization/gdnsq/layers"),
                      ("src.aux", "src/aux")]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(root, sub)]
            sys.modules[name] = m
    from src.quantization.gdnsq import gdnsq
    from src.quantization.gdnsq.layers import gdnsq_act, gdnsq_conv2d, gdnsq_linear
    from src.quantization.gdnsq.gdnsq_utils import QNMethod
    from src.aux.types import QScheme
    return types.SimpleNamespace(gdnsq=gdnsq, gdnsq_act=gdnsq_act, gdnsq_conv2d=gdnsq_conv2d,
                                 gdnsq_linear=gdnsq_linear, QNMethod=QNMethod, QScheme=QScheme,
                                 Quantizer=gdnsq.Quantizer, NoisyAct=gdnsq_act.NoisyAct,
                                 NoisyConv2d=gdnsq_conv2d.NoisyConv2d,
                                 NoisyLinear=gdnsq_linear.NoisyLinear, root=root, level="ops")


def load_full():
    """The whole reference package through its own __init__ files (Lightning & co. stubbed)."""
    root = ref_dir()
    stubs = install_stubs()
    _check_src_is_reference()
    # a previous load_ops() registered bare package modules: replace them by the real packages
    for name in [n for n in sys.modules if n == "src" or n.startswith("src.")]:
        m = sys.modules[name]
        if getattr(m, "__file__", None) is None and getattr(m, "__spec__", None) is None:
            del sys.modules[name]
    if root not in sys.path:
        sys.path.insert(0, root)
    import src.quantization as Q
    from src.quantization.quantizer import Quantizer as QuantizerFactory
    from src.quantization.gdnsq import gdnsq, gdnsq_quant, gdnsq_loss
    from src.quantization.gdnsq.layers import gdnsq_act, gdnsq_conv2d, gdnsq_linear
    from src.quantization.gdnsq.utils import model_helper, model_stats
    from src.quantization.gdnsq.calib import minmaxobserver, hooks
    from src.quantization.gdnsq.gdnsq_utils import QNMethod
    from src.aux.types import QScheme
    from src.models.cls.resnet import resnet_cifar
    from src.models.compose.vision.vision_cls_module import LVisionCls
    return types.SimpleNamespace(
        package=Q, QuantizerFactory=QuantizerFactory, GDNSQQuant=Q.GDNSQQuant, gdnsq=gdnsq,
        gdnsq_quant=gdnsq_quant, gdnsq_loss=gdnsq_loss, gdnsq_act=gdnsq_act,
        gdnsq_conv2d=gdnsq_conv2d, gdnsq_linear=gdnsq_linear, model_helper=model_helper,
        model_stats=model_stats, minmaxobserver=minmaxobserver, hooks=hooks, QNMethod=QNMethod,
        QScheme=QScheme, Quantizer=gdnsq.Quantizer, NoisyAct=gdnsq_act.NoisyAct,
        NoisyConv2d=gdnsq_conv2d.NoisyConv2d, NoisyLinear=gdnsq_linear.NoisyLinear,
        ModelHelper=model_helper.ModelHelper, PotentialLoss=gdnsq_loss.PotentialLoss,
        resnet_cifar=resnet_cifar, LVisionCls=LVisionCls, root=root, stubs=stubs, level="full")


def load_yaml_quantization(config_name: str):
    """The `quantization:` section of one of the reference's YAML configs as the attribute tree
    `GDNSQQuant` reads (the pydantic loader of src/config pulls the data/callback zoo; only
    this section concerns the path)."""
    import yaml
    path = os.path.join(ref_dir(), "config", config_name)
    with open(path) as f:
        cfg = yaml.safe_load(f)
    q = dict(cfg["quantization"])
    p = dict(q.get("params") or {})
    p.setdefault("distillation", False)
    p.setdefault("distillation_loss", "Cross-Entropy")
    p.setdefault("distillation_teacher", None)
    p.setdefault("qnmethod", "AEWGS")
    q["params"] = types.SimpleNamespace(**p)
    q.setdefault("excluded_layers", [])
    q.setdefault("freeze_batchnorm", False)
    q.setdefault("fuse_batchnorm", False)
    q.setdefault("quantize_bias", False)
    return types.SimpleNamespace(quantization=types.SimpleNamespace(**q), raw=cfg)


class FixedNoise:
    """Replace torch.randint_like (the reference's noise source, gdnsq.py:54) by a seeded
    {0,1} generator and record every draw, so a product run can be fed the same noise."""

    def __init__(self, seed, device="cpu"):
        self.gen = torch.Generator(device="cpu").manual_seed(seed)
        self.draws = []
        self._orig = torch.randint_like

    def __enter__(self):
        def fake(t, high, **kw):
            assert high == 2
            d = torch.randint(0, 2, t.shape, generator=self.gen).to(dtype=t.dtype, device=t.device)
            self.draws.append(d)
            return d.clone()
        torch.randint_like = fake
        return self

    def __exit__(self, *a):
        torch.randint_like = self._orig
