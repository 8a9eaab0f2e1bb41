"""Throughput of the DMMA block kernels (kb_block_gram / kb_block_apply) against the HBM roofline.

    python tools/bench_block.py [--rows 16777216] [--reps 20] [--quick]

Operands are larger than L2 (16.8 M rows x 16 columns = 2.1 GB each); CUDA events on the launching
stream; bytes are algorithmic (every operand element once).  FP64 flops are reported beside the
bandwidth: at k = l = 16 the kernels need 2 (gram) / 1.33 (apply) flop per byte."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from krylov_b200.device import BlockOps  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 24)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--quick", action="store_true", help="one shape, 2 repetitions (ncu capture)")
    a = ap.parse_args()
    peak = 6454.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    n = a.rows
    bo = BlockOps()
    shapes = [(16, 16)] if a.quick else [(16, 16), (8, 8), (16, 4), (4, 16), (1, 16), (12, 7)]
    reps = 2 if a.quick else a.reps
    print(f"rows {n}, peak {peak:.1f} GB/s (MEASURED_PEAKS.json hbm_gbs)")
    for k, l in shapes:
        X = torch.randn((n, k), dtype=torch.float64, device="cuda")
        Y = torch.randn((n, l), dtype=torch.float64, device="cuda")
        C = torch.randn((k, l), dtype=torch.float64, device="cuda")
        G = torch.zeros((k, l), dtype=torch.float64, device="cuda")
        Z = torch.empty((n, l), dtype=torch.float64, device="cuda")
        ms = timed(lambda: bo.gram(X, Y, out=G), reps)
        by = 8.0 * n * (k + l)
        print(f"gram  k={k:2d} l={l:2d}: {ms:7.3f} ms  {by / ms / 1e6:7.0f} GB/s = {by / ms / 1e6 / peak:5.3f} of peak"
              f"  {2.0 * n * k * l / ms / 1e9:6.2f} TFLOP/s fp64")
        ms = timed(lambda: bo.apply(X, C, Y=Y, sign=-1, out=Z), reps)
        by = 8.0 * n * (k + 2 * l)
        print(f"apply k={k:2d} l={l:2d}: {ms:7.3f} ms  {by / ms / 1e6:7.0f} GB/s = {by / ms / 1e6 / peak:5.3f} of peak"
              f"  {2.0 * n * k * l / ms / 1e9:6.2f} TFLOP/s fp64   (Z = Y - X C)")
        ref = timed(lambda: torch.matmul(X.t(), Y, out=G), max(2, reps // 4))
        print(f"      (cuBLAS X^T Y beside it: {ref:7.3f} ms)")
        del X, Y, Z


if __name__ == "__main__":
    main()
