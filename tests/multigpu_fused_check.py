"""Run under torchrun with N >= 2 GPUs: spp-sharded render on every rank, then the fused peer-memory
reduce + resolve kernel (buffers mapped through CUDA IPC) against the oracle's resolve of the host-side
sum of all ranks' buffers. Prints 'FUSED_OK' on rank 0. Used by tests/test_gpu_parity.py.
With --exchange: the same through rt_exchange_setup / rt_exchange_resolve, where the ranks order themselves with flags in
peer-mapped device memory (no barrier, no collective between render and resolve), three frames in a row with a reset in
between; prints 'EXCHANGE_OK'. Used by tests/test_gpu_round2.py."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracer_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rtb200  # noqa: E402
from oracle_py import Oracle  # noqa: E402


def exchange_main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h = 320, 200
    objs = np.load(os.path.join(ROOT, "tests", "golden", "bundled_scenes.npz"))["Scene3_indirect"]
    t = rtb200.PathTracer(local)
    t.set_scene(objs); t.set_camera(rtb200.default_camera())
    t.set_params(rtb200.default_params(width=w, height=h, mode=0, max_bounces=8, seed_lo=11, seed_hi=22))
    t.set_shard(rank, world)
    t.reset_accumulation()

    def gather(which):
        mine = torch.tensor(list(t.ipc_export(which)), dtype=torch.uint8, device="cuda")
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        return [bytes(x.cpu().tolist()) for x in allh]
    acc_h, srf_h, flg_h = gather(0), gather(1), gather(2)
    accum = [None if r == rank else t.ipc_open(acc_h[r]) for r in range(world)]
    flags = [None if r == rank else t.ipc_open(flg_h[r]) for r in range(world)]
    dst = None if rank == 0 else t.ipc_open(srf_h[0])
    t.exchange_setup(rank, world, accum, flags, dst)
    dist.barrier()                                       # setup only: every rank has mapped every buffer
    ok = True
    orc = Oracle()
    for frame, spp in enumerate((10, 3, 17)):
        t.reset_accumulation()
        if rank == world - 1 and frame == 1:
            time.sleep(0.25)                             # a late rank: the others wait for its signal ON THE DEVICE
        t.render_spp(spp)
        t.exchange_resolve(spp)                          # no host synchronisation, no collective
        surf = t.read_surface() if rank == 0 else None   # rank 0: ordered after every rank's slice by the exchange itself
        t.sync()
        mine, n = t.read_accum()
        parts = [torch.empty(h, w, 4, device="cuda") for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(mine).cuda())
        if rank == 0:
            total = np.zeros((h, w, 4), np.float32)
            for p in parts:
                total = total + p.cpu().numpy()
            ok = ok and np.array_equal(surf, orc.resolve_argb8(total, spp))
        dist.barrier()
    if rank == 0:
        print("EXCHANGE_OK" if ok else "EXCHANGE_MISMATCH", flush=True)
    t.close()
    dist.destroy_process_group()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h, spp = 320, 200, 10
    objs = np.load(os.path.join(ROOT, "tests", "golden", "bundled_scenes.npz"))["Scene3_indirect"]
    t = rtb200.PathTracer(local)
    t.set_scene(objs); t.set_camera(rtb200.default_camera())
    t.set_params(rtb200.default_params(width=w, height=h, mode=0, max_bounces=8, seed_lo=11, seed_hi=22))
    t.set_shard(rank, world)
    t.reset_accumulation()
    t.render_spp(spp)
    t.sync()

    def gather(which):
        mine = torch.tensor(list(t.ipc_export(which)), dtype=torch.uint8, device="cuda")
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        return [bytes(x.cpu().tolist()) for x in allh]
    acc_h, srf_h = gather(0), gather(1)
    ptrs = [t.accum_device_ptr() if r == rank else t.ipc_open(acc_h[r]) for r in range(world)]
    dst = t.argb_device_ptr() if rank == 0 else t.ipc_open(srf_h[0])
    dist.barrier()
    px = w * h
    first = px * rank // world
    t.resolve_fused(ptrs, spp, first, px * (rank + 1) // world - first, dst)
    t.sync()
    dist.barrier()
    mine, n = t.read_accum()
    parts = [torch.empty(h, w, 4, device="cuda") for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(mine).cuda())
    if rank == 0:
        total = np.zeros((h, w, 4), np.float32)
        for p in parts:
            total = total + p.cpu().numpy()
        ok = np.array_equal(t.read_surface(), Oracle().resolve_argb8(total, spp))
        # GPU-count invariance: the same global samples on one context
        one = rtb200.PathTracer(local)
        one.set_scene(objs); one.set_camera(rtb200.default_camera())
        one.set_params(rtb200.default_params(width=w, height=h, mode=0, max_bounces=8, seed_lo=11, seed_hi=22))
        one.reset_accumulation(); one.render_spp(spp)
        ok = ok and np.allclose(one.read_accum()[0], total, rtol=1e-6, atol=1e-6)
        print("FUSED_OK" if ok else "FUSED_MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    if "--exchange" in sys.argv:
        exchange_main()
    else:
        main()
