"""bench.py's contract on the CPU: the five BASELINE configurations are selectable, and the reference arm
(`--impl reference`: the CPU oracle port timed on the host cores) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_configs_are_the_baseline_configs():
    import bench
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert set(bench.CONFIGS) == {'cfg1', 'cfg2', 'cfg3', 'cfg4', 'cfg5'} and len(base['configs']) == 5
    bench.select_config('cfg4')
    assert bench.IMAGE[:2] == (128, 128) and bench.SAVP_HPARAMS['sequence_length'] == 16
    bench.select_config('cfg2')
    assert bench.IMAGE == (64, 64, 3) and bench.PER_GPU_BATCH == 16 and bench.SAVP_HPARAMS['sequence_length'] == 12
    assert bench.SAVP_HPARAMS['context_frames'] == 2


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '3', '--warmup', '1'],
                       capture_output=True, text=True, timeout=600, env=env)
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert r.returncode == 0 and len(lines) == 1, r.stderr[-500:]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['metric'].startswith('frames/sec SAVP 64x64') and d['n_gpus'] == 1 and d['data'] == 'synthetic'
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and abs(cb['value'] - d['value']) < 1e-9 and 'training steps' in cb['sample']
    assert d['e2e'] == dict(value=d['value'], unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d['config']['workload'] and d['ms_per_step'] > 0
