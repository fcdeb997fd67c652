"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, `--precision auto`
resolves per net, and the f8 numerics model stays where the design says (3-4 % of the single-pass fp16 error)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu_batch', '2',
                          '--num_steps', '3'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data',
              'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in line, k
    assert line['impl'] == 'reference' and line['value'] > 0 and line['higher_is_better'] is True and line['vs_baseline'] is None
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1 and line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == dict(value=line['value'], unit=line['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in line['config'] and 'model' not in line['config']


def test_precision_auto_resolves_per_net(monkeypatch):
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for net, want, fmin in (('cifar10', 'fp16f8', 0), ('imagenet64', 'fp16f8', 0), ('ffhq', 'fp16f8', 256), ('sd15', bench.PRECISION_FOR['sd15'], 0)):
        monkeypatch.setattr(sys, 'argv', ['bench.py', '--net', net])
        a = bench.parse()
        assert (a.precision, a.precision_requested, a.f8_min_channels) == (want, 'auto', fmin)
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--net', 'ffhq', '--precision', 'fp16f8'])
    a = bench.parse()
    assert a.precision == 'fp16f8' and a.f8_min_channels == 0          # an explicit precision keeps f8_min_channels as given


def test_f8_operand_model_error_budget():
    """CPU emulation of the operand formats of the f8 GEMM mode inside the oracle (tests/study_fp8_corrections.py): the denoiser error
    is a few percent of the single-pass fp16 error and well inside 1e-3 on the de-zeroed reduced-size nets."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import study_fp8_corrections as St
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    for name in ('tiny_song', 'tiny_adm'):
        r = {k: max(v) for k, v in St.run(name, batch=2, sigmas=(80.0, 0.5)).items()}
        assert r['fp16x3'] < 2e-5 and r['fp16+f8'] < 2e-4
        assert r['fp16+f8'] < 0.08 * r['fp16'], r
