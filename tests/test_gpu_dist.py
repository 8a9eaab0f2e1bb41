"""-m gpu multi-GPU test: 2 ranks under torchrun (NCCL over NVLink) must
reproduce the single-GPU CG/MINRES/GMRES histories on the same global problem.
Skipped on a box with one GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                    reason="needs 2 GPUs")
def test_two_rank_solvers_match_single_gpu(tmp_path):
    out = tmp_path / "dist.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29617",
           os.path.join(ROOT, "tools", "dist_check.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    for name, d in res.items():
        assert d["steps_equal"], (name, d)
        assert d["hist_rel"] <= 1e-9, (name, d)
        assert d["sol_rel"] <= 1e-10, (name, d)
        if name.startswith("fused_cg_"):
            assert d["fused_path_used"], (name, d)
            assert all(d["solve_success"]) and d["solve_final_rel"] <= 1e-10, (name, d)
