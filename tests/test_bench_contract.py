"""bench.py's output contract, checked on CPU through the reference arm (the sm_100a arm needs a GPU): stdout carries exactly
one JSON line even when libraries print to file descriptor 1, and the line has the keys the driver reads."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:400]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    # the reference's own classes are timed where /root/reference exists (here), the oracle port on the GPU box
    from oracle import ref_import
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_import.reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"] and d["config"]["bench_config"] == 2
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_stdout_is_reserved_for_the_json_line():
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench._claim_stdout(); "
            "os.write(1, b'NCCL version banner\\n'); print('python print'); bench._emit({'ok': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert r.stdout == '{"ok": 1}\n'
    assert "NCCL version banner" in r.stderr and "python print" in r.stderr
