"""bench.py's reference arm (the CPU restatement timed on host cores) runs without a GPU and
prints ONE JSON line with the keys the driver reads; the B200 arm refuses to run without CUDA
instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    res = _run('--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-seqs', '2')
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'tracked_frames_per_sec' and d['unit'] == 'frames/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['gpu_launches'] == 0
    assert d['value'] > 0 and d['steps'] == 1 and d['warmup'] == 1 and d['n_gpus'] == 1
    assert 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == dict(value=d['value'], unit=d['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_b200_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    res = _run('--steps', '1', '--warmup', '1', '--no-cpu-baseline')
    assert res.returncode != 0
    assert not any(ln.startswith('{') for ln in res.stdout.splitlines())      # no bench line from a CPU fallback


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """The driver launches the reference arm like the B200 arm (torchrun, N ranks): rank 0 alone
    measures and prints, the other ranks exit 0 without work."""
    import socket
    with socket.socket() as sock:                        # a free rendezvous port
        sock.bind(('127.0.0.1', 0))
        port = sock.getsockname()[1]
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'bench.py'),
                          '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1', '--cpu-seqs', '2'],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['value'] > 0
