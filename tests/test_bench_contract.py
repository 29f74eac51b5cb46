"""`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on a one-utterance sample: the JSON line
carries the contract's keys with the reference arm's meanings.  CPU only; the GPU arm's line is produced on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torchrun exports: bench.py must not inherit it for the CPU arm
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-batch", "1"], capture_output=True, text=True, timeout=580, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 1 and j["steps"] == 1
    assert j["metric"].startswith("mel-frames/sec training") and j["unit"] == "frames/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["vs_baseline"] is None
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == j["value"]
    assert cb["cores"] > 1 or os.cpu_count() == 1, cb     # OMP_NUM_THREADS=1 from the launcher was not inherited
    assert j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]


@pytest.mark.timeout(600)
def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """N > 1: the driver launches the reference arm with torchrun like the GPU arm; rank 0 alone measures and prints,
    the other ranks exit 0 without work, and the launcher's OMP_NUM_THREADS=1 does not reach the measurement."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-batch", "1"],
                       capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["value"] > 0
    assert j["cpu_baseline"]["cores"] > 1 or os.cpu_count() == 1
