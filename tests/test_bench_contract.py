"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (--impl reference)
prints one JSON line with the agreed keys on rank 0 and nothing on the other ranks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'A',
                           '--steps', '1', '--warmup', '1'], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == '256x256 crops/sec' and d['unit'] == 'crops/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1 and d['warmup'] >= 3     # bench.py never warms up less than 3 steps
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_other_ranks_are_silent():
    r = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith('{')]
