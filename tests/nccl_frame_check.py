"""Launched by tests/test_multi_gpu_nccl.py under torchrun (one rank per GPU): every rank renders its screen tiles, NCCL gathers
the tiled rgba8 buffers on rank 0, rtb_untile lays the frame out, and rank 0 compares it — and the shared host frame written by
rtb_present_host on every rank — with the frame it rendered alone.  Exits non-zero on any difference."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from igx_raytracing_b200 import rtb
    from conftest import synthetic_sky
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for kind in ("niels", "soup"):
        if kind == "niels":
            scene, limits, w, h, samples, eye = rtb.niels_scene(0.3), dict(), 333, 187, 2, (6, 5, 12)
        else:
            n = 100_000
            scene = dict(triangles=rtb.gen_soup(n, 0xB200), lights=rtb.niels_scene()["lights"][:32],
                         materials=rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0),
                         material_indices=np.zeros(n, np.uint32), info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))
            limits, w, h, samples, eye = dict(max_triangles=n), 640, 360, 1, (0.0, 0.0, 13.9)
        ctx = rtb.Context(device=local, **limits)
        stream = torch.cuda.Stream(device=local)
        torch.cuda.set_stream(stream)
        ctx.set_stream(stream.cuda_stream)
        ctx.resize(w, h, samples)
        ctx.upload_scene(scene, synthetic_sky())
        ctx.build_accel(rtb.ACCEL_BVH)
        cam = rtb.pack_camera(w, h, eye=eye)

        def one_frame():
            ctx.upload(rtb.BUF_CAMERA, cam)
            ctx.upload(rtb.BUF_SEED, rtb.make_seed((3.0, 9.0)))
            ctx.dispatch(rtb.PASS_FRAME)

        single = None
        if rank == 0:
            one_frame()
            single = ctx.readback(rtb.TGT_RGBA8).copy()
        ctx.set_option(rtb.OPT_TILE_COUNT, world)
        ctx.set_option(rtb.OPT_TILE_RANK, rank)
        one_frame()
        ptr, nbytes = ctx.device_ptr(rtb.TGT_RGBA8_TILED)
        slots = nbytes // 4

        class Arr:
            __cuda_array_interface__ = {"shape": (slots,), "typestr": "<i4", "data": (ptr, False), "version": 2, "strides": None}
        tiled = torch.as_tensor(Arr(), device=f"cuda:{local}")
        gathered = torch.empty(world * slots, dtype=torch.int32, device=f"cuda:{local}") if rank == 0 else None
        dist.gather(tiled, list(gathered.split(slots)) if rank == 0 else None, dst=0)
        # the shared host frame: every rank writes its own tiles
        from multiprocessing import shared_memory
        names = [None]
        shm = None
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=w * h * 4)
            names[0] = shm.name
        dist.broadcast_object_list(names, src=0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=names[0])
            try:   # attaching registers the segment with this process's resource tracker too (CPython < 3.13); only rank 0 owns it
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        host = np.frombuffer(shm.buf, dtype=np.uint32, count=w * h)
        assert int(torch.cuda.cudart().cudaHostRegister(host.ctypes.data, w * h * 4, 3)) == 0
        ctx.present_host(host.ctypes.data)
        staging = tiled.clone()
        side = torch.cuda.Stream(device=local)
        side.wait_stream(stream)
        ctx.present_host(host.ctypes.data, staging.data_ptr(), side.cuda_stream)   # the overlapped form: a copy of the tiles, a second stream
        side.synchronize()
        ctx.sync()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            ctx.untile(gathered.data_ptr(), world, slots, 0)
            got = ctx.readback(rtb.TGT_RGBA8)
            bad_gather = int((got != single).sum())
            bad_host = int((host.reshape(h, w) != single.reshape(h, w)).sum())
            print(f"{kind}: {world} ranks, {w}x{h}: gathered frame differs in {bad_gather} pixels, shared host frame in {bad_host}", flush=True)
            ok = ok and bad_gather == 0 and bad_host == 0 and bool(single.any())
        dist.barrier()
        torch.cuda.cudart().cudaHostUnregister(host.ctypes.data)
        del host
        shm.close()
        if rank == 0:
            shm.unlink()
        ctx.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
