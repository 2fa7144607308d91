"""Multi-GPU sharding invariance on real hardware (run under torchrun, one rank
per GPU; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py

Every rank shards one plane by row bands, exchanges halo rows over NCCL, runs
the band through morsi_cuda_apply_band_device and compares its rows bit for
bit with the same rows of the whole image processed on its own GPU."""
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = M.lib()
    check(L.morsi_cuda_init(local))
    ok = True
    w, h = 2048, 3001
    for element, op, kind in [("disk15", "tophat", 0), ("disk7", "closing", 0), ("cross", "gradient", 2),
                              ("disk5", "median", 0), ("dysk4", "oscillation", 2)]:
        e = M.parse_element(element)
        job = shard.BandJob(L, M.OPS.index(op), e, w, h, rank, world, dist, torch, seed=4, dist_kind=kind)
        job.step()
        torch.cuda.synchronize()
        band = job.y.cpu().numpy()
        # the whole image on this GPU
        full = torch.empty((h, w), dtype=torch.float32, device="cuda")
        out = torch.empty_like(full)
        check(L.morsi_cuda_synth(full.data_ptr(), w, h, 0, 0, 4, kind, job.stream))
        check(L.morsi_cuda_apply_device(M.OPS.index(op), job.e_p, full.data_ptr(), out.data_ptr(), w, h, 1, job.stream))
        torch.cuda.synchronize()
        want = out[job.plan.b0:job.plan.b1].cpu().numpy()
        nan = np.isnan(want)
        same = np.array_equal(np.isnan(band), nan) and \
            np.array_equal(band.view(np.uint32)[~nan], want.view(np.uint32)[~nan])
        print(f"rank {rank}/{world} {element} {op}: rows [{job.plan.b0},{job.plan.b1}) "
              f"{'bit-identical to the 1-GPU result' if same else 'MISMATCH'}", flush=True)
        ok &= same
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(1 if t.item() else 0)


if __name__ == "__main__":
    main()
