"""CPU: the N > 1 path's host logic with world_size-2 (and 3) gloo process groups: the C-ABI's shard
arithmetic (rt_shard_range) gives every global sample index to exactly one rank, and shard renders
summed with one all-reduce equal the single-rank image (here rendered by the CPU oracle standing in
for the kernels; the -m gpu tests check the kernels' side of the same contract)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rtb200
from conftest import GOLD, SEED


def test_shard_ranges_partition_the_samples():
    for world in (1, 2, 3, 4, 8):
        for spp in (0, 1, 5, 8, 1024, 1027):
            for start in (0, 17):
                got = []
                for r in range(world):
                    first, count = rtb200.shard_range(spp, r, world, start)
                    got.extend(range(first, first + count))
                assert got == list(range(start, start + spp))
    with pytest.raises(rtb200.RtError):
        rtb200.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_py import Oracle, OrcCamera
    orc = Oracle()
    objs = np.load(os.path.join(GOLD, "bundled_scenes.npz"))["Scene2"]
    cam = OrcCamera(); cam.right[0] = 1; cam.up[1] = 1; cam.forward[2] = 1; cam.fov_deg = 55
    w, h, spp = 48, 36, 7
    p = orc.default_params(width=w, height=h, max_bounces=8, mode=0, seed_lo=SEED[0], seed_hi=SEED[1])
    total = torch.zeros(h, w, 3)
    next_sample = 0
    for call_spp in (spp, 4):                                   # two rt_render_spp calls
        first, count = rtb200.shard_range(call_spp, rank, world, next_sample)
        part, _, _ = orc.render(objs, cam, p, first, count, threads=1)
        total += torch.from_numpy(part)
        next_sample += call_spp
    dist.all_reduce(total)                                      # the path's one exchange step
    if rank == 0:
        np.save(out_path, total.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_spp_sharded_sum_equals_single_rank(tmp_path, oracle, world):
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    from oracle_py import OrcCamera
    objs = np.load(os.path.join(GOLD, "bundled_scenes.npz"))["Scene2"]
    cam = OrcCamera(); cam.right[0] = 1; cam.up[1] = 1; cam.forward[2] = 1; cam.fov_deg = 55
    p = oracle.default_params(width=48, height=36, max_bounces=8, mode=0, seed_lo=SEED[0], seed_hi=SEED[1])
    want, _, _ = oracle.render(objs, cam, p, 0, 11)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)         # same samples, different float summation order
