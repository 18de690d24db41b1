"""bench.py's contract pieces that can be checked without a GPU: stdout carries nothing but the one JSON line (libraries and
argparse write to stderr), the product arm refuses to run without a CUDA device (no CPU fallback), and the reference arm's
non-zero ranks exit quietly."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(REPO / "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=e)


def test_help_and_chatter_stay_off_stdout():
    r = run(["--help"])
    assert r.returncode == 0 and r.stdout == "" and "--impl" in r.stderr


def test_product_arm_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and r.stdout == ""
    assert "no CPU fallback" in r.stderr


def test_reference_arm_other_ranks_exit_without_work():
    r = run(["--impl", "reference", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=120)
    assert r.returncode == 0 and r.stdout == ""
