"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the
row-band-sharded result must be bit-identical to the one-GPU result.
  * one process per GPU under torchrun (tests/mgpu_check.py): halo rows pushed
    between PROCESSES through CUDA IPC mappings by libmorsi_cuda;
  * one process, N devices (morsi_cuda_apply_sharded): peer access.
The torchrun log is kept in gpurun_out/mgpu_check.log (copied to profiles/)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import imscript_b200 as M
from oracle import oracle
from tests.test_gpu_parity import assert_same

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        return M.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")


@needs2
def test_sharded_processes_bit_identical():
    n = 2 if _ngpu() < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "mgpu_check.log"), "w") as f:
        f.write(f"$ {' '.join(cmd)}\nexit code {r.returncode}\n{r.stdout}\n--- stderr ---\n{r.stderr[-4000:]}\n")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout
    assert r.stdout.count("bit-identical to the 1-GPU result") >= 7 * n


@needs2
@pytest.mark.parametrize("name,op,iters", [("disk15", "tophat", 1), ("disk5", "erosion", 3), ("cross", "gradient", 1),
                                           ("disk4.2", "median", 1)])
def test_sharded_one_process_bit_identical(name, op, iters):
    o = oracle()
    ndev = min(_ngpu(), 4)
    h, w = 1501, 1024
    x = M.synth_host(w, h, seed=12, dist=0)
    e = o.element(name)
    want = x
    for _ in range(iters):
        want = M.apply(op, e, want)                  # one device
    got = M.apply_sharded(op, e, x, ndev, iters)
    assert_same(got, want, f"apply_sharded {name} {op} x{iters} on {ndev} devices")
    crop = o.apply(op, e, x[:200]) if iters == 1 else None
    if crop is not None:
        assert_same(got[:160], crop[:160], "apply_sharded vs oracle (top rows)")


def test_sharded_single_rank_and_errors():
    """nranks = 1 runs without neighbours; bad arguments are refused"""
    import ctypes
    from imscript_b200 import shard
    from imscript_b200.binding import check
    L = M.lib()
    check(L.morsi_cuda_init(0))
    o = oracle()
    e = M.parse_element("disk7")
    w, h = 512, 300
    job = shard.ShardJob(L, M.OPS.index("tophat"), e, w, h, 0, 1, 0, None, seed=9)
    job.step(); job.step()
    job.sync()
    band = np.empty((h, w), np.float32)
    check(L.morsi_cuda_memcpy_d2h(band.ctypes.data, job.out_ptr(), band.nbytes, job.stream))
    job.sync()
    assert_same(band, o.apply("tophat", o.element("disk7"), M.synth_host(w, h, seed=9)), "single-rank shard")
    # an operation that needs more halo than the shard holds
    big = M.parse_element("disk15")
    assert L.morsi_shard_apply(job.s, M.OPS.index("tophat"), big.ctypes.data_as(M.binding._i32p), 0, 2) == 1
    job.destroy()
    s = ctypes.c_void_p()
    assert L.morsi_shard_create(ctypes.byref(s), 0, 0, 4, 64, 40, 28, 2) == 1      # bands shorter than the halo
    assert L.morsi_shard_create(ctypes.byref(s), 0, 2, 2, 64, 400, 8, 2) == 1      # rank out of range


@needs2
def test_host_entry_point_on_two_devices_large_smem_kernels(monkeypatch):
    """MORSI_CUDA_DEVICES=2 with elements whose kernels opt in to > 48 KB of dynamic shared memory
    (hrec60: k_line, dysk12: k_tiled): the opt-in is per device (round-1 advice)"""
    o = oracle()
    x = np.stack([M.synth_host(700, 300, plane=p, seed=5) for p in range(2)])
    want = {}
    for name, op in [("hrec60", "dilation"), ("dysk12", "erosion"), ("disk5", "rank")]:
        want[(name, op)] = M.apply(op, o.element(name), x)
    monkeypatch.setenv("MORSI_CUDA_DEVICES", "2")
    for (name, op), y in want.items():
        assert_same(M.apply(op, o.element(name), x), y, f"2 devices {name} {op}")
    one = M.synth_host(700, 900, seed=6)
    e = o.element("hrec60")
    monkeypatch.delenv("MORSI_CUDA_DEVICES")
    y1 = M.apply("opening", e, one)
    monkeypatch.setenv("MORSI_CUDA_DEVICES", "2")
    assert_same(M.apply("opening", e, one), y1, "2 devices, one plane in row bands, hrec60 opening")
