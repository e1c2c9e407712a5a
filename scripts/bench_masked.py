"""The masked (cross-validation) ALS iteration at scale: iterations/s of c_ard_nmf's loop (reference src/singlet.cpp:1091-1152:
predict_mask both ways + scale + cor, mse_test every few iterations) on a synthetic slice, through the C ABI's sharded
fit (csrc/multi.cu sgl_fit_*, masked = 1 -> layout A: cells sharded for the H update, genes for the W update).
Runs on one GPU or under torchrun on N GPUs (NCCL). Prints one JSON line on rank 0.

    python scripts/bench_masked.py [m n density k] [--steps K] [--warmup W]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", nargs="*", default=["30000", "100000", "0.05", "32"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--inv-density", type=int, default=20)
    args = ap.parse_args()
    m, n, dens, k = int(args.shape[0]), int(args.shape[1]), float(args.shape[2]), int(args.shape[3])
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as dist

    from singlet_b200 import synth
    from singlet_b200.multi import RankComm, RankFit
    from singlet_b200.sharded import CudaBackend, shard_bounds

    torch.cuda.set_device(local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = CudaBackend(local)
    table = synth.values_table(m, dens)
    c0, c1, _ = shard_bounds(n, world, rank)
    g0, g1, _ = shard_bounds(m, world, rank)
    # layout A: this rank's cells (all genes) and this rank's genes over ALL cells
    A_sh = be.synth(m, n, dens, synth.DATA_SEED, 0, c0, c1 - c0, table)
    At_sh = be.synth(m, n, dens, synth.DATA_SEED, 1, g0, g1 - g0, table)
    be.synchronize()
    comm = RankComm(be._h, local, world, rank, group)
    t0 = time.perf_counter()
    fit = RankFit(comm, A_sh, At_sh, n, k, synth.w_init(k, m), masked=True, seed=4321, inv_density=args.inv_density)
    be.synchronize()
    t_masks = time.perf_counter() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        fit.iterate(0.01, 0.01, 0.0, 0.0)
    barrier()
    be.profile(True)
    be.profile_read()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        tol = fit.iterate(0.01, 0.01, 0.0, 0.0)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    prof = be.profile_read()
    be.profile(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mse = fit.test_mse()
    e1.record()
    barrier()
    ms_mse = e0.elapsed_time(e1)
    t = torch.tensor([ms, ms_mse], dtype=torch.float64, device=be.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fit.close()
    if rank == 0:
        ms, ms_mse = float(t[0]), float(t[1])
        print(json.dumps({"workload": f"masked ALS (c_ard_nmf loop) synthetic {m} x {n}, {dens:.0%}, k={k}, 1/{args.inv_density} held out",
                          "n_gpus": world, "steps": args.steps, "ms_per_iteration": ms / args.steps, "iterations_per_s": 1000.0 * args.steps / ms,
                          "mse_test_ms": ms_mse, "test_mse": mse, "tol": tol, "mask_build_s": t_masks,
                          "breakdown_ms_per_iteration_rank0": {kk: v[0] / args.steps for kk, v in prof.items()}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
