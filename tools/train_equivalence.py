"""System-level parity over a training run: the LIVE reference's own QAT pipeline (oracle/_ref:
its in-tree ResNet-20, LVisionCls, GDNSQQuant.quantize, calibration, patched training_step,
ModelHelper, PotentialLoss) trained for 40 steps
  (a) on the reference's own layers (its ATen op chain, eager), and
  (b) on this repo's layers swapped in per INTEGRATION.md §B (the sm_100a kernels),
same init, same batches, same GPU, deterministic cuDNN, TF32 off, deterministic estimator (LSQ for
weights and — overriding reference quirk 1 on both sides — for activations).
Bit-exact forward + bit-exact input gradients => the trajectories coincide until fp32
summation-order noise in the (tiny) scale gradients is amplified by training.

    python tools/train_equivalence.py > gpurun_out/r02_training_equivalence.txt
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as RH  # noqa: E402
from oracle import ref_loader  # noqa: E402

torch.backends.cudnn.deterministic = True
torch.backends.cudnn.benchmark = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
steps, B = 40, 128
ref = ref_loader.load_full()
g = torch.Generator(device=dev).manual_seed(0)
xs = [torch.randn(B, 3, 32, 32, device=dev, generator=g) for _ in range(4)]
ts = [torch.randint(0, 10, (B,), device=dev, generator=g) for _ in range(4)]
torch.manual_seed(123)
base = ref.resnet_cifar.resnet20_cifar10(num_classes=10)
cfg = RH.make_cfg(ref, act_bit=4, weight_bit=4, qscheme=1, qnmethod="LSQ", excluded_layers=["conv1", "linear"])


def run(swapped):
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_act import NoisyAct
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_conv2d import NoisyConv2d
    from mhaq_b200.quantization.gdnsq.layers.gdnsq_linear import NoisyLinear
    from mhaq_b200.quantization.gdnsq.gdnsq_utils import QNMethod
    import contextlib
    ctx = RH.swapped_layers(ref, NoisyAct, NoisyConv2d, NoisyLinear) if swapped else contextlib.nullcontext()
    with ctx:
        lm = RH.build_lmodule(ref, copy.deepcopy(base), 10, lr=1e-3).to(dev)
        q = RH.quantize(ref, lm, cfg).to(dev)
        RH.calibrate(ref, q, xs[0], act_bits=8, weight_bits=8, device=dev)
        for m in q.model.modules():            # deterministic activations on both sides
            if isinstance(m, (NoisyAct, ref.NoisyAct)):
                m.Q.qnmethod = QNMethod.LSQ if swapped else ref.QNMethod.LSQ
        q.wrapped_criterion.t = 0.0
        opt = torch.optim.SGD([p for p in q.parameters() if p.requires_grad], lr=2e-3, momentum=0.9)
        q.train(); q.wrapped_criterion.train()
        losses = []
        for i in range(steps):
            loss = q.training_step((xs[i % 4], ts[i % 4]), i)
            loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
            losses.append(float(loss.detach()))
        return losses, {n: p.detach().clone() for n, p in q.model.named_parameters()}


la, pa = run(True)
lb, pb = run(False)
print("step   loss(reference pipeline on the sm_100a layers)   loss(reference pipeline, own layers)   rel.diff")
for i in range(steps):
    if i < 10 or i % 5 == 4:
        print(f"{i:4d}   {la[i]:.7f}                                        {lb[i]:.7f}                            {abs(la[i]-lb[i])/abs(lb[i]):.2e}")
num = sum(float((pa[n].float() - pb[n].float()).double().pow(2).sum()) for n in pa)
den = sum(float(pb[n].float().double().pow(2).sum()) for n in pa)
print(f"relative L2 distance between the two parameter sets after {steps} steps: {(num / den) ** 0.5:.2e}")
print(f"loss went {la[0]:.4f} -> {la[-1]:.4f} (sm_100a layers), {lb[0]:.4f} -> {lb[-1]:.4f} (reference layers)")
