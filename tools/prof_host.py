"""Host-side (Python) profile of one QAT training step — where does the CPU time go?"""
import cProfile, pstats, sys, os, io, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhaq_b200 import harness
model = sys.argv[1] if len(sys.argv) > 1 else "resnet20"
method = sys.argv[2] if len(sys.argv) > 2 else "STE"
dev = torch.device("cuda")
side, classes, B = (32, 100, 256) if model == "resnet20" else (224, 1000, 128)
x = torch.randn(B, 3, side, side, device=dev); t = torch.randint(0, classes, (B,), device=dev)
q = harness.build_qat(model, dev, qnmethod=method, distillation=True, num_classes=classes, calib_batch=x[:64])
opt = q.configure_optimizers(); q.train(); q.wrapped_criterion.train()
def step():
    loss = q.training_step((x, t), 0); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(5): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host time/step {1e3*(t1-t0)/10:.2f} ms, incl. drain {1e3*(t2-t0)/10:.2f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
