"""Per-tensor clamped STE backward at 2^29 / 2^30 elements: single-launch flat kernel vs the
streaming kernel + finalize (MHAQ_FQ_FLAT_MAX_LOG2 decides).  CUDA-event timing of eager calls
(kernels of 1-2 ms: launch overhead is noise)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhaq_b200 import ops

for log2n in (29, 30):
    n = 1 << log2n
    x = torch.randn(n, device="cuda"); go = torch.randn(n, device="cuda")
    b = torch.tensor([-2.0], device="cuda"); s = torch.tensor([0.25], device="cuda"); hi = b + 4.0 - s
    L = ops._Launch(x, s, b, b, hi)
    f = lambda: ops._backward_impl(go, x, L, 0, False, None, True, philox=(1, 2))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"2^{log2n}: {ms*1e3:.1f} us  {12*n/ms/1e6:.0f} GB/s  single_launch={ops.lib.mhaq_fq_bwd_single_launch(1, n, 1, 0, 0)}")
    del x, go
