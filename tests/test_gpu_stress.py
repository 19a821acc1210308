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


@pytest.mark.gpu
@pytest.mark.parametrize("cases,seed,round_size", [(12, 3, 1), (8, 4, 8)])
def test_randomised_sweep_through_the_plugin_layer(cases, seed, round_size):
    """tools/stress_harness.py: random estimator / sampler (uniform, PROSAC, NAPSAC over the grid or the k nearest neighbours) / SPRT /
    LO through ransac_b200/usac/usac_harness --both. The fused Ransac::run() must equal the oracle's rounds of K, the one-hypothesis-at-
    a-time loop over the plug-in classes the oracle's sequential loop (= the compiled reference's semantics), both after the refit:
    iterations, inliers, inlier list, model bits. (By hand: 140 runs, profiles/r2_stress_harness.txt.)"""
    harness_dir = os.path.join(ROOT, "ransac_b200", "usac")
    subprocess.check_call(["make", "-s", "-C", harness_dir])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_harness.py"), str(cases), str(seed), str(round_size), "napsac"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and ", 0 mismatches" in r.stdout, r.stdout[-4000:]
