"""CPU: bench.py's output contract, checked on its real main() with the GPU work replaced by stand-ins (tests/_mock_bench.py):
exactly ONE JSON line on stdout carrying the driver's keys; an optional block that raises is reported inside the line; an
optional block that stalls past --extras-budget does not take the headline with it (the line is still printed, once, and
the process exits 0)."""
import json
import os
import subprocess
import sys
from conftest import ROOT

KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
        'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline', 'train', 'configs', 'sweep'}


def run(env=None, extra=()):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', '_mock_bench.py')] + list(extra), capture_output=True, text=True,
                       env=e, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_one_json_line_with_the_contract_keys():
    d = run()
    assert KEYS <= set(d), KEYS - set(d)
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['n_gpus'] == 1
    assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(d['e2e'])
    assert d['e2e']['h2d_bytes_per_step'] == 3 * 600 * 1000 * 4
    assert d['train'] == {'graph': True} and d['configs']['3_fpn'] == {'test': 1, 'train': 2} and 'extras' not in d


def test_failing_optional_block_is_reported_not_fatal():
    d = run({'MOCK_FAIL': '1'})
    assert d['train'] == {'failed': 'boom'} and d['configs']['2_deformable_faster'] == {'images_per_sec': 700}


def test_deadline_keeps_the_headline():
    d = run({'MOCK_HANG': '1'}, ['--extras-budget', '6'])
    assert 'stopped at the 6 s deadline' in d['extras']
    assert d['value'] > 0 and d['configs']['3_fpn'] == {'test': 1}          # the part reported before the stall is kept


def test_sticky_device_error_in_an_optional_block_still_prints_the_line():
    d = run({'MOCK_STICKY': '1'})
    assert d['value'] > 0 and d['train'] == {'eager': True}                # the part reported before the error is kept
    assert d['extras'].startswith('aborted: ') and 'device unusable' in d['extras']
    assert d['configs']['2_deformable_faster'] == {'images_per_sec': 700} and d['configs']['3_fpn'] is None
