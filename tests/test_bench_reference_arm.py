"""``bench.py --impl reference``: the CPU arm the driver runs next to the GPU arm. Its JSON line carries the bench
contract's keys plus ``impl`` / ``cpu_baseline`` / a zero-copy ``e2e``; with more than one host core the sample runs as
concurrent single-threaded replicas (one per core) and ``cores`` says how many; ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

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


def test_reference_arm_line_and_replicas():
    one = json.loads(_run(['--cpu-procs', '1']).splitlines()[-1])
    two = json.loads(_run(['--cpu-procs', '2']).splitlines()[-1])
    for line, cores in ((one, 1), (two, 2)):
        for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                    'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
            assert key in line, key
        assert line['impl'] == 'reference' and line['metric'] == 'INS s/timestep' and line['unit'] == 's'
        assert line['higher_is_better'] is False and line['value'] > 0
        cb = line['cpu_baseline']
        assert cb['kind'] == 'port' and cb['cores'] == cores and cb['value'] == line['value']
        assert line['e2e'] == {'value': line['value'], 'unit': 's', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'replicas' in two['cpu_baseline']['sample'] and 'replicas' not in one['cpu_baseline']['sample']
    assert one['config'] == two['config']


def test_reference_arm_other_ranks_stay_silent():
    assert _run(['--gpus', '2'], env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == ''
