"""Host-side helpers of bench.py that do not need a GPU."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_usable_cores_respects_affinity_and_is_bounded():
    b = _bench()
    n = b.usable_cores()
    assert 1 <= n <= 32
    if hasattr(os, "sched_getaffinity"):
        assert n <= len(os.sched_getaffinity(0))


def test_cpu_samples_are_valid_configs():
    """Every bounded sample must be a config the discriminator accepts (side >= 128, divisible by 32) and ordered largest first."""
    b = _bench()
    costs = [side * side * t for side, t in b.CPU_SAMPLES]
    assert costs == sorted(costs, reverse=True)
    for side, t in b.CPU_SAMPLES:
        assert side >= 128 and side % 32 == 0 and t >= 1


def test_reference_arm_other_ranks_exit_silently():
    """Under torchrun only rank 0 runs the reference arm; the other ranks print nothing and exit 0 (bench contract)."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_flop_model_matches_design():
    """one step = B * [2 (F_G + 6 F_D) + K (3 F_G + 3 F_D)] with F_G = 521.4 GF, F_D = 35.7 GF."""
    b = _bench()
    got = b.flop_step(16, 1)
    assert abs(got / 1e12 - 50.3) < 0.5, got
