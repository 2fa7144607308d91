"""Multi-GPU sharding invariance on real hardware, run under torchrun with one
rank per GPU (spawned by tests/test_gpu_multi.py when >= 2 GPUs are visible):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py

Every rank shards one plane by row bands through the C API (morsi_shard_*: halo
rows pushed over NVLink by libmorsi_cuda itself, CUDA IPC between the ranks),
and compares its rows bit for bit with the same rows of the whole image
processed on its own GPU -- for several steps in a row (the credit / ready
protocol must hold when buffers are reused), for the host-band entry point, and
for an iterated operation (op applied 3 times, bands resident in between)."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import imscript_b200 as M                      # noqa: E402
from imscript_b200 import shard                # noqa: E402
from imscript_b200.binding import check        # noqa: E402


def same(a, b):
    nan = np.isnan(b)
    return np.array_equal(np.isnan(a), nan) and np.array_equal(a.view(np.uint32)[~nan], b.view(np.uint32)[~nan])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = M.lib()
    check(L.morsi_cuda_init(local))
    ok = True
    w, h = 2048, 3001

    def whole_image(op, e_p, kind, times=1):
        n = w * h * 4
        a, b = M.DeviceBuffer(n), M.DeviceBuffer(n)
        check(L.morsi_cuda_synth(a.ptr, w, h, 0, 0, 4, kind, None))
        for _ in range(times):
            check(L.morsi_cuda_apply_device(M.OPS.index(op), e_p, a.ptr, b.ptr, w, h, 1, None))
            a, b = b, a
        out = a.to_host((h, w))
        a.free(); b.free()
        return out

    for element, op, kind in [("disk15", "tophat", 0), ("disk7", "closing", 0), ("cross", "gradient", 2),
                              ("disk5", "median", 0), ("dysk4", "oscillation", 2), ("hrec9", "dilation", 0)]:
        e = M.parse_element(element)
        job = shard.ShardJob(L, M.OPS.index(op), e, w, h, rank, world, local, dist, seed=4, dist_kind=kind)
        p = job.plan
        want = whole_image(op, job.e_p, kind)[p.b0:p.b1]
        for overlap_steps in range(3):                     # three steps in a row: buffers 0, 1, 0
            job.step()
        job.sync()
        band = np.empty((p.rows_own, w), np.float32)
        check(L.morsi_cuda_memcpy_d2h(band.ctypes.data, job.out_ptr(), band.nbytes, job.stream))
        job.sync()
        good = same(band, want)
        # host rows in, host rows out
        hx = np.ascontiguousarray(M.synth_host(w, p.rows_own, row0=p.b0, seed=4, dist=kind))
        hy = np.empty_like(hx)
        check(L.morsi_shard_apply_host(job.s, job.op, job.e_p, hx.ctypes.data, hy.ctypes.data))
        good_host = same(hy, want)
        print(f"rank {rank}/{world} {element} {op}: rows [{p.b0},{p.b1}) "
              f"{'bit-identical to the 1-GPU result' if good else 'MISMATCH'}; host band "
              f"{'bit-identical' if good_host else 'MISMATCH'}; halo bytes/step {job.halo_bytes()}", flush=True)
        ok &= good and good_host
        job.destroy()

    # iterated operation: 3 x disk5 opening^... erosion, bands resident, ping-pong buffers 0 <-> 1
    e = M.parse_element("disk5")
    job = shard.ShardJob(L, M.OPS.index("erosion"), e, w, h, rank, world, local, dist, seed=4, dist_kind=0, nbuf=2)
    p = job.plan
    for it in range(3):
        check(L.morsi_shard_apply(job.s, job.op, job.e_p, it & 1, (it & 1) ^ 1))
    job.sync()
    band = np.empty((p.rows_own, w), np.float32)
    check(L.morsi_cuda_memcpy_d2h(band.ctypes.data, job.buffer_ptr(1) + p.own_offset * w * 4, band.nbytes, job.stream))
    job.sync()
    good = same(band, whole_image("erosion", job.e_p, 0, times=3)[p.b0:p.b1])
    print(f"rank {rank}/{world} disk5 erosion x3 (resident bands): "
          f"{'bit-identical to the 1-GPU result' if good else 'MISMATCH'}", flush=True)
    ok &= good
    job.destroy()

    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(1 if t.item() else 0)


if __name__ == "__main__":
    main()
