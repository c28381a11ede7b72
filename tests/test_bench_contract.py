"""bench.py --impl reference on the CPU: the JSON line of the reference arm (the driver parses it next to the b200
arm's line) carries the keys of the measurement contract, with the arm's own metric / unit / workload description."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                        '--workload', 'cpu_ref'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'queries/s' and line['higher_is_better'] is True
    assert line['metric'].startswith('VOGNet fwd queries/sec') and line['value'] > 0
    assert line['steps'] == 1 and line['warmup'] >= 3 and line['n_gpus'] == 1
    # the unmodified reference (baseline/_ref or /root/reference) when installed, the oracle port otherwise
    from oracle import ref_harness
    assert line['cpu_baseline']['kind'] == ('reference' if ref_harness.reference_available() else 'port')
    assert line['cpu_baseline']['cores'] >= 1 and line['reference_ranks'] == 1 and line['reference_n_queries'] == 1
    assert line['cpu_baseline']['value'] == line['value'] == line['e2e']['value']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    cfg = line['config']
    assert cfg['workload'] == 'cpu_ref' and cfg['obj_attn'] == [1, 50, 512] and cfg['mul_attn'] == [10, 25, 768]
    assert line['vs_baseline'] is None and line['data'] == 'synthetic' and line['gpu_launches'] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''
