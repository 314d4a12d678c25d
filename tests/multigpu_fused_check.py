"""Run under torchrun with N >= 2 GPUs: spp-sharded render on every rank, then the fused peer-memory
reduce + resolve kernel (buffers mapped through CUDA IPC) against the oracle's resolve of the host-side
sum of all ranks' buffers. Prints 'FUSED_OK' on rank 0. Used by tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracer_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rtb200  # noqa: E402
from oracle_py import Oracle  # noqa: E402


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
    main()
