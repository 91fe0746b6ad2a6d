"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port) prints exactly one JSON line with the
agreed keys; the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--small', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'images/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['metric'] == 'images/sec at 1024^2 bs8 ViT-B attn-shift' and 'workload' in d['config']
    assert set(d['cpu_baseline']) >= {'value', 'unit', 'cores', 'kind', 'sample'} and d['cpu_baseline']['kind'] == 'port'
    assert d['e2e'] == dict(value=d['value'], unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--small', '--gpus', '2', '--steps', '1'],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--small', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == ''      # fails loudly, prints no number
