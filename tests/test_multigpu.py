"""The partitioned paths on REAL GPUs, one process per GPU over NCCL (skipped with fewer than two
visible GPUs): block-column distributed Cholesky + block-column gradient == single-GPU
ExactGP._update / loglikelihood(True) (pygp/inference/exact.py:50-55,118-143) at 1e-10 / 1e-8,
sharded predict and sharded batched evaluation == one rank.  The host logic of the same paths
runs on CPU over gloo in tests/test_sharding.py and tests/test_distchol.py."""

import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(nproc, args, timeout):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, 'tools', 'dist_check.py')]
    out = subprocess.run(cmd + [str(a) for a in args], cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith('{')]


@pytest.mark.timeout(900)
@pytest.mark.parametrize('n,nb,kern', [(4608, 256, 'se8'), (8192 + 100, 512, 'c5')])
def test_two_ranks_equal_one(n, nb, kern):
    if _ngpu() < 2:
        pytest.skip('needs two GPUs')
    recs = {r['check']: r for r in _torchrun(2, [n, nb, kern], 800)}
    ev = recs['distributed_eval']
    assert ev['world'] == 2 and ev['lZ_rel_err'] <= 1e-10 and ev['dlZ_rel_err'] <= 1e-8
    assert ev['mu_err'] <= 1e-9 and ev['s2_err'] <= 1e-9
    assert recs['sharded_predict']['max_err'] <= 1e-12
    assert recs['sharded_batched_loglike']['rel_err_first4'] <= 1e-10


@pytest.mark.timeout(1800)
def test_c5_n65536_all_gpus():
    """BASELINE configs[4] at full size: SE + Periodic, N = 65536, every visible GPU."""
    g = _ngpu()
    if g < 2:
        pytest.skip('needs two or more GPUs')
    recs = {r['check']: r for r in _torchrun(g, [65536, 512, 'c5'], 1700)}
    ev = recs['distributed_eval']
    assert ev['lZ_rel_err'] <= 1e-10 and ev['dlZ_rel_err'] <= 1e-8
