"""The multi-GPU path is data-parallel replicas over contiguous image shards with no collective
on the data path.  This checks the sharding arithmetic and the optional logit
gather with two gloo ranks on the CPU: each rank runs ITS shard of a synthetic batch (through the
oracle, standing in for a GPU replica), logits are all-gathered, and the result must equal the
single-process run image for image."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, n, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT / "vision-transformer-opencl_b200"))
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_py as O
    import vit_b200 as V
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = V.shard_range(n, world, rank)
    w = V.synth_weights(224, 42)
    imgs = V.synth_images(hi - lo, 224, 7, first_index=lo)      # image i depends only on (seed, i)
    logits = O.forward(w, imgs, 224, n_threads=2) if hi > lo else np.zeros((0, 1000), np.float32)
    per = -(-n // world)
    mine = torch.zeros(per, 1000)
    mine[: hi - lo] = torch.from_numpy(logits)
    gathered = [torch.zeros(per, 1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    dist.barrier()
    if rank == 0:
        full = torch.cat([g[: V.shard_range(n, world, r)[1] - V.shard_range(n, world, r)[0]] for r, g in enumerate(gathered)])
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_single_process(tmp_path, vit, oracle, weights224):
    import torch.multiprocessing as mp
    n, world = 3, 2   # ragged: shards of 2 and 1
    mp.spawn(_worker, args=(world, 29500 + os.getpid() % 400, n, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    ref = oracle.forward(weights224, vit.synth_images(n, 224, 7), 224)
    assert got.shape == (n, 1000)
    assert np.array_equal(got, ref)
