"""bench.py's reference arm on a small stream (CPU only): the JSON line carries the keys the driver's contract names, and the
arm times the oracle (kind "port") on the same metric / unit / workload definition as the chord_b200 arm.  The chord_b200 arm
itself needs a GPU (it fails loudly without one - there is no CPU fallback - which is checked here too)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)


def test_reference_arm_json_line():
    p = _run(['--impl', 'reference', '--steps', '2', '--warmup', '1', '--meshlets', '4000'])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'Mtri/s software-rasterized' and d['unit'] == 'Mtri/s' and d['higher_is_better'] is True
    for k in ('value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['steps'] == 2 and d['warmup'] == 1 and d['value'] > 0 and d['vs_baseline'] is None and d['data'] == 'synthetic'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'Mtri/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert d['config']['triangles_per_step'] == 4000 * 128


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 without printing."""
    p = _run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', '--meshlets', '2000'], env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert p.returncode == 0 and p.stdout.strip() == '', (p.returncode, p.stdout, p.stderr[-500:])


def test_product_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('GPU present')
    p = _run(['--steps', '1', '--warmup', '0', '--meshlets', '2000', '--no-frame', '--no-cpu-baseline'])
    assert p.returncode != 0 and not any(l.startswith('{') for l in p.stdout.splitlines())
