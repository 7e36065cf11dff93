"""bench.py driver contract, the parts that can be checked without a GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_baseline_md():
    sys.path.insert(0, ROOT)
    import bench
    init_b, iter_b = bench.algorithmic_bytes(512, 640, 4, 32)
    assert init_b == 10158080 and iter_b == 51200000
    assert abs((init_b + 4 * iter_b) / 1e6 - 214.96) < 0.01
    init5, iter5 = bench.algorithmic_bytes(1056, 1920, 7, 48)          # config 5
    assert abs((init5 + 4 * iter5) / 1e6 - 2027.5) < 0.1


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "refs/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("reference-views/sec at 640x512")
    from oracle import reference_arm as RA
    # the unmodified reference when tools/install_ref.py has installed it (build() does, wherever /root/reference exists),
    # the oracle port otherwise
    assert d["cpu_baseline"]["kind"] == ("reference" if RA.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "refs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_install_is_byte_identical():
    """baseline/_ref (git-ignored, travels with gpurun) holds unmodified copies: sha256 manifest intact, and equal to
    /root/reference where that exists (the build container)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import install_ref
    state = install_ref.install(verbose=False)
    if state == "unavailable":
        pytest.skip("no /root/reference and no baseline/_ref on this machine")
    assert install_ref.installed()
    if os.path.isdir(install_ref.REF):
        for f in install_ref.FILES:
            assert install_ref._sha(os.path.join(install_ref.REF, f)) == install_ref._sha(os.path.join(install_ref.DST, f)), f
    tracked = subprocess.run(["git", "ls-files", "baseline"], cwd=ROOT, capture_output=True, text=True).stdout.split()
    assert not [t for t in tracked if t.startswith("baseline/_ref")], "reference sources must not be committed"


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_arm_fails_loudly_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "needs a GPU" in r.stderr
