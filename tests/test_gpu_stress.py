"""Randomised parity sweep (tools/stress_parity.py): 150 fits with random estimator / sampler / SPRT / LO / round size / problem size /
seed, each compared with the CPU oracle field by field, plus the ordered inlier list and the refit loop on both sides of the
multi-launch switch. A second seed is run by hand for profiles/r2_stress_parity.txt (400 fits, 0 mismatches)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_randomised_parity_sweep():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_parity.py"), "150", "3"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=900)
    assert r.returncode == 0 and ", 0 mismatches" in r.stdout, r.stdout[-4000:]


@pytest.mark.gpu
def test_randomised_parity_sweep_fallback_paths():
    """The same sweep through the fallback paths: LO as one CTA (taken when the candidate lists of the speculative waves would not fit)
    and SPRT batches one problem at a time."""
    env = dict(os.environ, USAC_GPU_LO_SEQ="1", USAC_GPU_SPRT_BATCH="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_parity.py"), "60", "11"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=900, env=env)
    assert r.returncode == 0 and ", 0 mismatches" in r.stdout, r.stdout[-4000:]
