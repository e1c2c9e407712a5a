"""CV rank sweep on pbmc3k with the (rank, replicate) grid dealt over N GPUs (one process per GPU, no data-path
collective):  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_cv_multi.py
Prints one JSON line on rank 0 (sweep seconds: first call and repeated; equality with the single-GPU sweep)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from singlet_b200 import api  # noqa: E402
from singlet_b200.datasets import get_pbmc3k_data, log_normalize  # noqa: E402
from singlet_b200.sharded import distributed_cross_validate_nmf  # noqa: E402

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
A = log_normalize(get_pbmc3k_data())
ranks = list(range(2, 31))
times = []
h = api.Handle(local)
for _ in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    df = distributed_cross_validate_nmf(A, ranks, n_replicates=3, seed=123, device=local, handle=h)
    if world > 1:
        dist.barrier()
    times.append(time.perf_counter() - t0)
if rank == 0:
    api.set_seed(123)
    ref = api.cross_validate_nmf(A, ranks, n_replicates=3, verbose=0, handle=api.Handle(local))
    sys.stderr.flush()
    os.write(1, (json.dumps({"n_gpus": world, "fits": len(ranks) * 3, "cv_sweep_first_s": round(times[0], 3),
                             "cv_sweep_repeat_s": round(min(times[1:]), 3), "equals_single_gpu_sweep": bool(df.equals(ref)),
                             "best_rank": int(api.GetBestRank(df))}) + "\n").encode())
if world > 1:
    dist.destroy_process_group()
