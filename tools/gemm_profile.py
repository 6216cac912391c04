"""Profiling / tuning target for the DMMA GEMM: a few launches at given shapes on
device buffers (torch = allocator only).  usage: gemm_profile.py M N K [reps]
PGP_GEMM_VARIANT=0|1 selects the warp layout (2x4 warps of 64x32 | 4x4 of 32x32)."""
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygp_b200 import _lib  # noqa: E402

args = [int(a) for a in sys.argv[1:]]
if len(args) == 1:
    args = [args[0]]*3
if len(args) == 2:
    args = [args[0]]*3 + [args[1]]
m, n, k = args[:3]
reps = args[3] if len(args) > 3 else 4
ctx, L = _lib.context(), _lib.lib()
a = torch.randn(m, k, dtype=torch.float64, device='cuda')
b = torch.randn(n, k, dtype=torch.float64, device='cuda')
c = torch.zeros(m, n, dtype=torch.float64, device='cuda')
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(ctx.stream)
best = 1e30
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    _lib.check(ctx, L.pgp_dev_gemm_nt(ctx.handle, m, n, k, -1.0, a.data_ptr(), k, b.data_ptr(), k, 1.0, c.data_ptr(), n, 0))
    e1.record(stream)
    ctx.sync()
    best = min(best, e0.elapsed_time(e1))
print('gemm_nt variant=%s %dx%dx%d: %.3f ms  %.2f TFLOP/s' % (os.environ.get('PGP_GEMM_VARIANT', '0'), m, n, k, best, 2.0*m*n*k/best/1e9))
