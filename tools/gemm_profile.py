"""Profiling target for `ncu --set full -k regex:gemm_nt`: a few launches of the
DMMA GEMM at one square size on device buffers (torch = allocator only)."""
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygp_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx, L = _lib.context(), _lib.lib()
a = torch.randn(n, n, dtype=torch.float64, device='cuda')
b = torch.randn(n, n, dtype=torch.float64, device='cuda')
c = torch.zeros(n, n, dtype=torch.float64, device='cuda')
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(ctx.stream)
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    _lib.check(ctx, L.pgp_dev_gemm_nt(ctx.handle, n, n, n, -1.0, a.data_ptr(), n, b.data_ptr(), n, 1.0, c.data_ptr(), n, 0))
    e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1)
    print('gemm_nt %d^3: %.3f ms  %.2f TFLOP/s' % (n, ms, 2*n**3/ms/1e9))
