"""Writes profiles/fp64_peak.json on a GPU box: the FP64 roofline denominators the driver's
MEASURED_PEAKS.json does not carry -- DMMA issue peak and DFMA peak (tools/fp64_peak.cu: register-operand
loops, best over 4..32 warps/SM) and cuBLAS DGEMM 8192^3 (torch.matmul, best of 10, CUDA events) --
with the clocks seen.  bench.py cites this file next to the cuBLAS DGEMM it re-measures in every run.
  python tools/fp64_peak.py            (after `make -C tools`)"""
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rec = json.loads(subprocess.check_output([os.path.join(ROOT, 'tools', 'bin', 'fp64_peak')],
                                             stderr=subprocess.DEVNULL).decode().strip().splitlines()[-1])
    m = 8192
    a = torch.randn(m, m, dtype=torch.float64, device='cuda')
    b = torch.randn(m, m, dtype=torch.float64, device='cuda')
    c = torch.empty(m, m, dtype=torch.float64, device='cuda')
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    rec['cublas_dgemm_8192_tflops'] = 2*m**3/(best*1e-3)/1e12
    # sustained: back to back for ~3 s
    t0 = time.perf_counter()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < 3.0:
        for _ in range(10):
            torch.matmul(a, b, out=c)
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    rec['cublas_dgemm_8192_tflops_sustained'] = n*2*m**3/(e0.elapsed_time(e1)*1e-3)/1e12
    q = subprocess.check_output(['nvidia-smi', '--query-gpu=name,clocks.sm,clocks.max.sm,power.draw', '--format=csv,noheader',
                                 '-i', '0']).decode().strip()
    rec['nvidia_smi'] = q
    rec['how'] = ('tools/fp64_peak.cu: mma.sync.m8n8k4.f64 / fma.f64 register-operand loops, best of 5 per warp count; '
                  'torch.matmul f64 8192^3 best of 10 (burst) and back to back for 3 s (sustained), CUDA events')
    rec['when'] = time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime())
    out = os.path.join(ROOT, 'gpurun_out' if os.path.isdir(os.path.join(ROOT, 'gpurun_out')) else 'profiles', 'fp64_peak.json')
    json.dump(rec, open(out, 'w'), indent=1)
    print(json.dumps(rec))


if __name__ == '__main__':
    main()
