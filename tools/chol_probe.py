"""Factorisation timing sweep (GPU box): pgp_dev_potrf at several N, with the
lookahead path on/off (PGP_CHOL_LOOKAHEAD), against cuSOLVER through torch."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pygp_b200 import _lib  # noqa: E402

ctx, L = _lib.context(), _lib.lib()
for n in [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 16384]:
    g = torch.randn(n, n + 64, dtype=torch.float64, device='cuda')
    K = g @ g.T/(n + 64) + torch.eye(n, dtype=torch.float64, device='cuda')
    del g
    F = K.clone()
    ref = torch.linalg.cholesky(K)
    best = 1e9
    for _ in range(3):
        F.copy_(K)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info = L.pgp_dev_potrf(ctx.handle, F.data_ptr(), n, n, 0)
        best = min(best, time.perf_counter() - t0)
    torch.cuda.synchronize()
    tc = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        torch.linalg.cholesky(K)
        torch.cuda.synchronize()
        tc = min(tc, time.perf_counter() - t0)
    err = float((torch.tril(F) - ref).abs().max())
    print(json.dumps({'n': n, 'lookahead': os.environ.get('PGP_CHOL_LOOKAHEAD', '1'), 'info': info, 'ours_ms': best*1e3,
                      'ours_tflops': n**3/3/best/1e12, 'cusolver_ms': tc*1e3, 'cusolver_tflops': n**3/3/tc/1e12,
                      'max_abs_err': err}), flush=True)
    del K, F, ref
    torch.cuda.empty_cache()
