"""bench.py contract, CPU side: the reference arm runs without a GPU (it times the reference's own classes on the host cores) and prints
ONE JSON line with the keys the driver reads; ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _run(env_extra):
    env = dict(os.environ, OMP_NUM_THREADS="1", **env_extra)   # torchrun exports OMP_NUM_THREADS=1: must not throttle the arm
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_json_line():
    lines = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ComA vertex-pairs/s" and d["unit"] == "vertex-pairs/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["n_gpus"] == 2 and d["steps"] == 1
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    from oracle import ref_loader
    # the UNMODIFIED reference classes when they are staged (oracle/_ref; always in the dev container), else the C port
    assert cb["kind"] == ("reference" if ref_loader.available() else "port") and cb["value"] == d["value"] and cb["sample"]
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert cb["cores"] == avail                                  # every available core despite OMP_NUM_THREADS=1
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
