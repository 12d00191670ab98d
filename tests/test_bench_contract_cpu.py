"""bench.py's CPU-runnable contract: the reference arm prints ONE JSON line with the agreed keys, and the GPU arm refuses to
run without a device (no silent fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    proc = run('--impl', 'reference', '--gpus', '1', '--steps', '2', '--warmup', '1')
    assert proc.returncode == 0, proc.stderr
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('acoustic frames/sec') and d['n_gpus'] == 1 and d['steps'] == 2
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['vs_baseline'] is None and d['scaling'] == 'weak'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_under_torchrun_env_only_rank0_prints():
    proc = run('--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1',
               env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert proc.returncode == 0 and proc.stdout.strip() == ''


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    proc = run('--gpus', '1', '--steps', '1', '--warmup', '1')
    assert proc.returncode != 0
    assert 'no CPU fallback' in (proc.stderr + proc.stdout)
