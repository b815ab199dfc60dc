"""Real multi-GPU runs (skipped on a single-GPU box): the NCCL sample-shard ring against the oracle, and
bench.py's partitioned arm at a small size, both launched the way the driver launches them (torchrun)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def torchrun(n, script, *args, port=29610):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script)] + list(args)
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)


@pytest.mark.skipif(NGPU < 2, reason="needs at least 2 GPUs")
def test_nccl_ring_blocks_match_the_oracle(built):
    n = 4 if NGPU >= 4 else 2
    p = torchrun(n, "scripts/ring_demo.py", "--check", port=29611)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["parity_vs_oracle"] is True and line["cells_computed"] == line["cells_strict_lower_triangle"]


@pytest.mark.skipif(NGPU < 2, reason="needs at least 2 GPUs")
def test_bench_partitioned_arm_small(built):
    n = 4 if NGPU >= 4 else 2
    p = torchrun(n, "bench.py", "--gpus", str(n), "--steps", "2", "--warmup", "3", "--samples", "1024", "--length", "600000",
                 "--no-cpu-baseline", port=29612)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == n and line["parity_vs_oracle"] is True and line["value"] > 0 and line["e2e"]["value"] > 0
