"""``bench.py --impl reference``: the CPU arm the driver runs next to the GPU arm. It MEASURES the CPU restatement on a
sample of the workload that finishes (value = the measured seconds per step of that sample, config = the sample's own
size), keeps the linear extrapolation to the full size as a labelled extra, and — when a GPU is present — times the GPU
path on the same sample in the same invocation (``gpu_same_config`` / ``measured_ratio``). Ranks other than 0 print
nothing."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    e = dict(os.environ)
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        e.pop(k, None)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '1', '--cpu-N', '4'] + extra, cwd=ROOT, env=e, capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout.strip()


@pytest.mark.parametrize('workload', ['ins3d_dim', 'ins2d'])
def test_reference_arm_line_is_a_measurement_of_its_own_config(workload):
    line = json.loads(_run(['--workload', workload]).splitlines()[-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'scaled_to_full_config'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['metric'] == 'INS s/timestep' and line['unit'] == 's'
    assert line['higher_is_better'] is False and line['value'] > 0
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] == 1 and cb['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': 's', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    # the line is about the sample it ran: its config names the sample's N, the value fits in the run's wall time
    assert line['config']['N'] == 4 and 'sample_of' in line['config']
    assert line['value'] * line['steps'] < 600
    assert 'EXTRAPOLATION' in line['scaled_to_full_config']['note']
    assert line['scaled_to_full_config']['value'] > line['value']
    assert 'gpu_same_config' not in line or 'error' in line['gpu_same_config'] or line['measured_ratio'] > 0


def test_reference_arm_other_ranks_stay_silent():
    assert _run(['--gpus', '2'], env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == ''


def test_cpu_sample_shrinks_when_many_steps_are_asked_for():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.cpu_sample_size('ins3d_dim', 5) == 8 and bench.cpu_sample_size('ins3d_dim', 20) == 6
    assert bench.cpu_sample_size('ins2d', 5) == 28 and bench.cpu_sample_size('ins2d', 20) == 20
    assert bench.cpu_sample_size('ins2d', 20, cpu_N=12) == 12
