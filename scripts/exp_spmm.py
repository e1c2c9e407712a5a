"""Developer probe: the SpMM right-hand-side product in both operand precisions on a slice of BASELINE configs[2]/[4].

    python scripts/exp_spmm.py [--cells 500000] [--genes 30000] [--k 32] [--density 0.05] [--reps 5]

Prints per launch: milliseconds (CUDA events inside the library, sgl_profile), algorithmic GB/s, and the difference
between the FP16-operand result and the FP32-operand result (max and RMS relative to the column's largest entry).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlet_b200 import _lib, synth  # noqa: E402
from singlet_b200.sharded import CudaBackend  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=500000)
    ap.add_argument("--genes", type=int, default=30000)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--density", type=float, default=0.05)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--modes", default="fp32,mixed16")
    args = ap.parse_args()
    m, n, k = args.genes, args.cells, args.k
    be = CudaBackend(0)
    table = synth.values_table(m, args.density)
    A = be.synth(m, n, args.density, synth.DATA_SEED, 0, 0, n, table)
    At = be.transpose(A)
    nnz = be.matrix_info(A)[2]
    kp = be.kp(k)
    g = torch.Generator(device="cuda").manual_seed(1)
    out = {"m": m, "n": n, "k": k, "nnz": nnz}
    for side, X, rows, cols in (("H-update (gather W)", A, m, n), ("W-update (gather H)", At, n, m)):
        F = torch.rand((rows, kp), generator=g, device="cuda", dtype=torch.float32)
        F[:, k:] = 0
        F = F * (torch.rand((rows, 1), generator=g, device="cuda") ** 4)  # rows of very different size
        F = F / F.sum(dim=0, keepdim=True).clamp_min(1e-30)               # like scale(): every factor sums to 1
        F[:, k:] = 0
        res = {}
        for mode in args.modes.split(","):
            _lib.check(be.lib.sgl_set_precision(be._h, {"fp32": 1, "mixed16": 0}[mode]))
            B = torch.zeros((cols, kp), device="cuda", dtype=torch.float32)
            be.rhs(X, F, k, B)  # builds the tile index / streams
            be.synchronize()
            be.profile(True)
            for _ in range(args.reps):
                be.rhs(X, F, k, B)
            prof = be.profile_read()
            be.profile(False)
            ms, cnt, byt = prof["spmm"]
            res[mode] = B
            line = {"side": side, "mode": mode, "ms_per_launch": ms / cnt, "alg_GBps": byt / cnt / (ms / cnt) / 1e6}
            print(json.dumps(line), flush=True)
            out[f"{side}/{mode}"] = line
        if "fp32" in res and "mixed16" in res:
            a, b = res["fp32"].double(), res["mixed16"].double()
            scale = a.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
            rel = ((a - b).abs() / scale)
            print(json.dumps({"side": side, "max_rel_to_colmax": float(rel.max()), "rms_rel": float((rel ** 2).mean().sqrt()),
                              "rel_fro": float((a - b).norm() / a.norm())}), flush=True)
    be.close()


if __name__ == "__main__":
    main()
