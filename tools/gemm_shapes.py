"""GEMM rate of the shapes the distributed update launches (M x nb x K updates of a block column inside the
N x N factor buffer): dense operands vs operands that are windows of a buffer with leading dimension N.
  python tools/gemm_shapes.py"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pygp_b200 import _lib  # noqa: E402


def main():
    ctx, L = _lib.context(), _lib.lib()
    N = 32768
    F = torch.randn(N, N, dtype=torch.float64, device='cuda')
    for (M, Nn, K) in [(16384, 1024, 2048), (16384, 1024, 1024), (16384, 512, 512), (16384, 512, 2048), (16384, 2048, 2048),
                       (16384, 4096, 2048), (8192, 1024, 2048)]:
        for layout in ('window_ldN', 'dense'):
            if layout == 'dense':
                A = torch.randn(M, K, dtype=torch.float64, device='cuda')
                C = torch.zeros(M, Nn, dtype=torch.float64, device='cuda')
                pa, lda, pb, ldb, pc, ldc = A.data_ptr(), K, A.data_ptr(), K, C.data_ptr(), Nn
            else:
                # C = F[r0:, c1:c1+Nn], A = F[r0:, 0:K], B = F[r0:r0+Nn, 0:K]  (as catch_up in dist.cu)
                r0, c1 = N - M, K
                pa, lda = F.data_ptr() + (r0*N)*8, N
                pb, ldb = pa, N
                pc, ldc = F.data_ptr() + (r0*N + c1)*8, N
            for tri in (0, 1):
                def run():
                    _lib.check(ctx, L.pgp_dev_gemm_nt(ctx.handle, M, Nn, K, -1.0, pa, lda, pb, ldb, 1.0, pc, ldc, tri))
                run()
                ctx.sync()
                best = 1e9
                for _ in range(3):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    run()
                    ctx.sync()
                    best = min(best, time.perf_counter() - t0)
                print(json.dumps({'M': M, 'N': Nn, 'K': K, 'layout': layout, 'tri': tri, 'ms': best*1e3,
                                  'tflops': 2.0*M*Nn*K/best/1e12}), flush=True)


if __name__ == '__main__':
    main()
