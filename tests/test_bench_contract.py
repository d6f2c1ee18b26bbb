"""bench.py's CPU legs (no GPU needed): the reference arm of every workload prints ONE JSON line with the contract's keys, the
launcher-safe process-group shutdown returns under gloo, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run_bench(*args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


@pytest.mark.parametrize("workload,metric,unit", [("predict", "generated FLAME frames/sec", "frames/s"),
                                                  ("prior", None, "samples/s"), ("train", "training clips/sec", "clips/s")])
def test_reference_arm_prints_the_contract_line(workload, metric, unit):
    r = run_bench("--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["data"] == "synthetic"
    assert d["unit"] == unit and (metric is None or d["metric"] == metric)
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    assert d["vs_baseline"] is None                       # BASELINE.md publishes no number for these metrics
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_under_a_launcher_runs_on_rank_0_only():
    """N > 1: rank 0 prints the line, the other ranks exit 0 without work (no process group is needed for the CPU arm)."""
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")], (r.stdout[-500:], r.stderr[-500:])


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run_bench("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_process_group_shutdown_returns_under_gloo(tmp_path):
    script = tmp_path / "sd.py"
    script.write_text(
        "import sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import torch, torch.distributed as dist\n"
        "import bench\n"
        "dist.init_process_group('gloo')\n"
        "x = torch.ones(4); dist.all_reduce(x)\n"
        "assert x.tolist() == [2.0] * 4\n"
        "bench.shutdown_process_group(dist.get_world_size())\n"
        "print('done', flush=True)\n")
    import socket
    with socket.socket() as sk:                       # a free rendezvous port (another job on the host may hold a fixed one)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count("done") == 2
