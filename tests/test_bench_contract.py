"""bench.py contract checks that need no GPU: the reference arm (the reference's own CPU
fast_sampler Session, oracle/_ref, or the C port) runs end to end at a tiny scale and prints the
one JSON line the driver parses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.002",
                        "--steps", "3", "--warmup", "3"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "sampled_and_gathered_minibatches_per_sec"
    assert line["unit"] == "batches/s" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["value"] > 0 and line["steps"] == 3 and line["warmup"] >= 3
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "batches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["config"]["workload"].startswith("ogbn-papers100M-shaped")


def test_reference_arm_distributed_sessions_at_two_gpus():
    """N > 1: rank 0 drains N distributed reference Sessions (book + cache), cores / N threads each."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--scale",
                        "0.002", "--steps", "4", "--warmup", "3"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference-distributed" and line["config"]["feature_partitions"] == 8
    assert line["e2e"]["value"] == line["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--scale",
                        "0.01", "--steps", "3", "--warmup", "3"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_tools_and_entry_points_compile():
    """Every helper script must at least byte-compile (they only run on the GPU box)."""
    import glob
    import py_compile
    paths = glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, "bench.py"),
                                                              os.path.join(ROOT, "__graft_entry__.py"),
                                                              os.path.join(ROOT, "tests", "multigpu_check.py")]
    assert len(paths) >= 8
    for p in paths:
        py_compile.compile(p, doraise=True)
