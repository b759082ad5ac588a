"""The N > 1 host logic under real process groups on CPU (gloo, world size 2): partitioning agreement between ranks, and the
reference arm of bench.py under torchrun (rank 0 works and prints, the other ranks exit 0 without work)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torchrun(port, *script_and_args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd + list(script_and_args), capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_slab_partitioning_agrees_across_two_ranks():
    r = torchrun(29533, os.path.join(ROOT, "tests", "slab_host_check.py"))
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["result"] == "ok" and d["world"] == 2 and len(d["ranges"]) == 2 and sum(d["n_own"]) == 600 * 120


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    r = torchrun(29534, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--cpu-columns", "20",
                 "--rows", "100", "--presteps", "5")
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
