#!/usr/bin/env python
"""bench.py -- morsi hot-path benchmark (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl ours|reference]

Prints ONE JSON line.  A "step" is one pass of the hot path over one batch of
synthetic input; the default workload is BASELINE.json configs[1] ("c2":
disk7 opening AND closing of a 4096x4096 float32 RGB image => 2 x 3 x 4096^2
samples per step per GPU).  `value` is whole-job Msample/s ("Mpixel/s" in the
reference's vocabulary: one pixel = one float32 sample of one plane) with the
input resident in HBM; `e2e` is the same metric through morsi_cuda_apply()
with pinned HOST buffers, copies inside the timed region.

Under torchrun (N>1) every rank processes its own frames (no data-path
collective, "weak") except for --workload c4, where one 40000x40000 plane is
row-band sharded and halo rows are exchanged between ranks every step
("strong") inside libmorsi_cuda (morsi_shard_*: peer stores over NVLink).  Every
line of every other workload also carries that C4 run at the same N as the
"sharded" sub-record, so a 1/2/4/8-GPU sweep of the default workload holds the
north-star strong-scaling curve.  torch is used for the process group (barriers,
max over ranks, the one-off gather of the 128-byte shard handles) only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (element, [ops], w, h, planes, seed, description)
    "c1": ("square", ["erosion"], 1024, 1024, 1, 1, "square erosion 1024x1024x1 (BASELINE configs[0])"),
    "c2": ("disk7", ["opening", "closing"], 4096, 4096, 3, 2,
           "disk7 opening+closing 4096x4096x3 (BASELINE configs[1])"),
    "c3": ("disk5", ["median"], 8192, 8192, 1, 3, "disk5 median 8192x8192x1 (BASELINE configs[2])"),
    "c4": ("disk15", ["tophat"], 40000, 40000, 1, 4,
           "disk15 tophat 40000x40000x1, row-band sharded (BASELINE configs[3])"),
    "c5": ("cross", ["gradient"], 1920, 1080, 3 * 64, 5,
           "cross gradient, 64 RGB 1920x1080 frames per GPU per step (chunk of BASELINE configs[4])"),
}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks under the benchmark's load (B200_PROFILING.md recipe).
    nvidia-smi needs ~100 ms to start and samples every 20 ms, while a timed
    region can be a few milliseconds: the sampler is started before the
    warm-up, and the caller keeps the SAME load running (untimed) after the
    timed region until a few samples have fallen inside the load window."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.t_load0 = self.t_load1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line))

    def load_begin(self):
        self.t_load0 = time.time()

    def samples_under_load(self):
        return sum(1 for t, _ in self.lines if self.t_load0 is not None and t >= self.t_load0 + 0.002)

    def stop(self):
        self.t_load1 = time.time()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for t, line in list(self.lines):
            if self.t_load0 is None or t < self.t_load0 + 0.002 or t > self.t_load1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                sm.append(float(f[2])); mx.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "window": "timed region + the same load continued untimed until >= 5 samples"}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return world, rank, local, dist, torch
    return 1, 0, 0, None, None


# ----------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ----------------------------------------------------------------------------
def synth_numpy(w, rows, row0=0, plane=0, seed=1):
    """morsi_synth_value (imscript_b200/csrc/common.cuh), distribution 0, restated with
    numpy so that the reference arm never maps libmorsi_cuda.so."""
    def mix(hh):
        hh = hh.astype(np.uint32)
        hh ^= hh >> np.uint32(16); hh = (hh * np.uint32(0x85EBCA6B)).astype(np.uint32)
        hh ^= hh >> np.uint32(13); hh = (hh * np.uint32(0xC2B2AE35)).astype(np.uint32)
        hh ^= hh >> np.uint32(16)
        return hh
    with np.errstate(over="ignore"):
        h0 = mix(np.array([(seed * 0x9E3779B1 + plane) & 0xFFFFFFFF], dtype=np.uint64).astype(np.uint32))
        r = (np.arange(row0, row0 + rows, dtype=np.uint64) * 0x27D4EB2F).astype(np.uint32)
        hr = mix(h0 ^ r)[:, None]
        c = (np.arange(w, dtype=np.uint64) * 0x165667B1).astype(np.uint32)[None, :]
        hc = mix(hr ^ c)
    return ((hc >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def reference_sample_shape(name, big):
    """A bounded crop of the workload: ~10-15 s of single-thread CPU work for the
    cpu_baseline leg (big), ~2-3 s per step for the --impl reference arm."""
    if big:
        return {"c1": (1024, 1024), "c2": (2048, 1536), "c3": (2048, 1536), "c4": (1536, 1024),
                "c5": (1920, 1080)}[name]
    return {"c1": (1024, 1024), "c2": (1024, 768), "c3": (1024, 768), "c4": (768, 512),
            "c5": (1920, 1080)}[name]


def run_reference_once(name, threads, seed_offset=0, big=False):
    """All `threads` host threads run the reference (oracle/_ref, else the oracle
    port) on their own crop, the reference's own parallelism doctrine
    (doc/misc/optimization.txt:43-49: several single-threaded programs at once).
    Returns (samples processed, seconds, kind, description)."""
    from oracle import oracle as get_oracle
    from oracle.oracle import Reference
    element, ops, w, h, planes, seed, _ = WORKLOADS[name]
    cw, ch = reference_sample_shape(name, big)
    o = get_oracle()
    e = o.element(element)
    kind = "port"
    impl = o
    if os.path.exists(Reference.path):
        from oracle import reference as get_ref
        impl, kind = get_ref(), "reference"
    crops = [synth_numpy(cw, ch, row0=(h - ch) // 2, plane=t, seed=seed + seed_offset) for t in range(threads)]

    def work(t):
        for op in ops:
            impl.apply(op, e, crops[t])        # ctypes releases the GIL

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return threads * cw * ch * len(ops), dt, kind, f"{threads} crop(s) of {cw}x{ch} x {len(ops)} op(s)"


def workload_config(name, world=1):
    element, ops, w, h, planes, seed, desc = WORKLOADS[name]
    sharded = name == "c4"
    samples = w * h * planes * len(ops) * (1 if sharded else world)
    return {"workload": desc, "element": element, "ops": ops, "samples_per_step": samples,
            "l2": "input+output per op exceed the 126 MB L2" if w * h * planes * 8 > 126e6
            else "working set fits L2 (launch-latency-bound config)"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        run_reference_once(name, threads)
    total, secs = 0, 0.0
    for k in range(args.steps):
        n, dt, kind, sample = run_reference_once(name, threads, seed_offset=k)
        total += n
        secs += dt
    value = total / secs / 1e6
    print(json.dumps({
        "impl": "reference", "metric": "morsi Mpixel/s", "value": value, "unit": "Mpixel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": "strong" if name == "c4" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": threads, "kind": kind,
                         "sample": sample + " per step, one crop per host thread (throughput does not depend on the "
                                            "image size: SURVEY 8d), the unmodified src/morsi.c compiled by oracle/Makefile"},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
class Env:
    """what every measurement leg needs"""

    def __init__(self, args):
        import imscript_b200 as M
        self.M, self.L = M, M.lib()
        self.check = M.binding.check
        self.ct = M.binding.ctypes
        self.world, self.rank, self.local, self.dist, self.torch = dist_setup(args.gpus)
        if M.device_count() < 1:
            raise SystemExit("bench.py: no CUDA device; libmorsi_cuda has no CPU fallback")
        self.check(self.L.morsi_cuda_init(self.local))
        self.args = args

    def barrier(self, stream=None):
        self.check(self.L.morsi_cuda_sync(stream))
        if self.dist is not None:
            self.torch.cuda.synchronize()
            self.dist.barrier()

    def max_over_ranks(self, v):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return float(t.item())

    def timed(self, step, stream, steps, warmup, sync=None):
        """W warm-up steps, barrier, EXACTLY `steps` steps between two CUDA events on the
        launching stream, barrier; max over ranks.  Returns (ms_per_step, launches, clocks)."""
        L, check, ct = self.L, self.check, self.ct
        sync = sync or (lambda: check(L.morsi_cuda_sync(stream)))
        ev = [self.M.binding._vp() for _ in range(2)]
        for x in ev:
            check(L.morsi_cuda_event_create(ct.byref(x)))
        sampler = ClockSampler(self.local)
        sampler.start()
        for _ in range(warmup):
            step()
        sync()
        self.barrier(stream)
        sampler.load_begin()
        L.morsi_cuda_launch_count_reset()
        check(L.morsi_cuda_event_record(ev[0], stream))
        for _ in range(steps):
            step()
        check(L.morsi_cuda_event_record(ev[1], stream))
        sync()
        launches = L.morsi_cuda_launch_count()
        ms = ct.c_float()
        check(L.morsi_cuda_event_elapsed_ms(ev[0], ev[1], ct.byref(ms)))
        self.barrier(stream)
        elapsed_ms = self.max_over_ranks(ms.value)
        launches = int(self.sum_over_ranks(launches))
        ms_per_step = elapsed_ms / steps
        # keep the same load running (untimed) until the clock sampler has seen it; under
        # torchrun every rank runs the same number of extra steps (a step may exchange halo
        # rows with its neighbours)
        if self.dist is not None:
            for _ in range(min(5000, int(200.0 / max(ms_per_step, 1e-3)) + 1)):
                step()
        else:
            t_more = time.time()
            while sampler.proc and sampler.samples_under_load() < 5 and time.time() - t_more < 1.5:
                for _ in range(max(1, steps // 4)):
                    step()
                sync()
        sync()
        self.barrier(stream)
        clocks = sampler.stop()
        for x in ev:
            L.morsi_cuda_event_destroy(x)
        return ms_per_step, launches, clocks

    def wall(self, step, steps):
        """host wall clock around synchronous steps, max over ranks -> seconds per step"""
        step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        return self.max_over_ranks(time.perf_counter() - t0) / steps


def pcie_ceiling(env, hx, hy, nbytes):
    """Aggregate pinned H2D + D2H rate of the box with every rank copying both ways at
    once (the e2e path can do no better): GB/s per direction summed over ranks."""
    L, check, ct, M = env.L, env.check, env.ct, env.M
    nb = min(nbytes, 1 << 30)
    d_a, d_b = M.DeviceBuffer(nb), M.DeviceBuffer(nb)
    s1, s2 = M.binding._vp(), M.binding._vp()
    check(L.morsi_cuda_stream_create(ct.byref(s1)))
    check(L.morsi_cuda_stream_create(ct.byref(s2)))
    reps = 3

    def both():
        check(L.morsi_cuda_memcpy_h2d(d_a.ptr, hx, nb, s1))
        check(L.morsi_cuda_memcpy_d2h(hy, d_b.ptr, nb, s2))
        check(L.morsi_cuda_sync(s1))
        check(L.morsi_cuda_sync(s2))
    sec = env.wall(both, reps)
    L.morsi_cuda_stream_destroy(s1)
    L.morsi_cuda_stream_destroy(s2)
    d_a.free(); d_b.free()
    return env.world * nb / sec / 1e9


def measure_cli(name, hx_array_fn):
    """Wall time of the drop-in command line itself on this workload's image (the path a user of the
    reference runs): `morsi ELEMENT OP in.npy out.npy`, malloc'd iio buffers, process start-up and
    CUDA context creation included.  Returns None when the CLI or a scratch directory is missing."""
    import shutil
    import tempfile
    element, ops, w, h, planes, seed, desc = WORKLOADS[name]
    cli = os.path.join(ROOT, "imscript_b200", "lib", "morsi")
    if not os.path.exists(cli) or w * h * planes * 4 > (1 << 30):
        return None
    tmp = tempfile.mkdtemp(prefix="morsi_bench_")
    try:
        fin, fout = os.path.join(tmp, "in.npy"), os.path.join(tmp, "out.npy")
        np.save(fin, hx_array_fn())                       # (h, w, planes): iio's pixel-interleaved NPY
        t0 = time.perf_counter()
        for op in ops:
            r = subprocess.run([cli, element, op, fin, fout], capture_output=True)
            if r.returncode != 0:
                return {"error": r.stderr.decode()[-200:]}
        dt = time.perf_counter() - t0
        nbytes = os.path.getsize(fin)
        return {"value": w * h * planes * len(ops) / dt / 1e6, "unit": "Mpixel/s", "seconds": dt,
                "file_bytes_in": nbytes, "what": "`morsi %s OP in.npy out.npy` for OP in %s: process start, iio read, "
                "upload / kernels / download, iio write (page cache, no disk)" % (element, ops)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def measure_planes(env, name):
    """plane / frame workloads: every rank its own planes, no data-path collective"""
    M, L, check, ct = env.M, env.L, env.check, env.ct
    args = env.args
    element, ops, w, h, planes, seed, desc = WORKLOADS[name]
    e = M.parse_element(element)
    e_p = e.ctypes.data_as(M.binding._i32p)
    opi = [M.OPS.index(o) for o in ops]
    n = w * h * planes
    d_x, d_y = M.DeviceBuffer(n * 4), M.DeviceBuffer(n * 4)
    for p in range(planes):
        check(L.morsi_cuda_synth(d_x.ptr + p * w * h * 4, w, h, 0, p + env.rank * planes, seed, 0, None))
    check(L.morsi_cuda_sync(None))

    def step():
        for o in opi:
            check(L.morsi_cuda_apply_device(o, e_p, d_x.ptr, d_y.ptr, w, h, planes, None))
    ms_per_step, launches, clocks = env.timed(step, None, args.steps, args.warmup)

    # ---- e2e: the public host-pointer call, pinned host buffers ----------
    hx, hy = M.binding._vp(), M.binding._vp()
    nbytes = n * 4
    check(L.morsi_cuda_host_alloc(ct.byref(hx), nbytes))
    check(L.morsi_cuda_host_alloc(ct.byref(hy), nbytes))
    check(L.morsi_cuda_memcpy_d2h(hx, d_x.ptr, nbytes, None))
    check(L.morsi_cuda_sync(None))
    d_x.free(); d_y.free()
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        for o in opi:
            check(L.morsi_cuda_apply(o, e_p, hx, hy, w, h, planes))
    sec = env.wall(e2e_step, e2e_steps)
    samples_total = n * len(ops) * env.world
    ceil_gbs = pcie_ceiling(env, hx, hy, nbytes)
    e2e = {"value": samples_total / sec / 1e6, "unit": "Mpixel/s",
           "h2d_bytes_per_step": nbytes * len(ops) * env.world, "d2h_bytes_per_step": nbytes * len(ops) * env.world,
           "steps": e2e_steps, "api": "morsi_cuda_apply (host pointers, pinned)",
           "ceiling": {"value": ceil_gbs / 4 * 1e3, "unit": "Mpixel/s",
                       "pcie_gbs_per_direction_all_ranks": ceil_gbs,
                       "how": "every rank copying 1 GiB pinned H2D and D2H at once; 4 B up + 4 B down per sample"}}
    e2e["frac_of_ceiling"] = e2e["value"] / e2e["ceiling"]["value"]
    if env.rank == 0 and env.world == 1 and not args.no_cpu:
        def as_vec():
            a = np.ctypeslib.as_array(ct.cast(hx, ct.POINTER(ct.c_float)), shape=(planes, h, w))
            return np.ascontiguousarray(np.moveaxis(a, 0, -1))
        e2e["cli"] = measure_cli(name, as_vec)
    L.morsi_cuda_host_free(hx)
    L.morsi_cuda_host_free(hy)
    return {"ms_per_step": ms_per_step, "launches": launches, "clocks": clocks, "e2e": e2e,
            "samples_total": samples_total, "scaling": "weak"}


def measure_sharded(env, steps, warmup, with_e2e=True):
    """C4: ONE 40000x40000 plane, row-band sharded over the ranks (strong scaling), halo rows
    pushed over NVLink inside libmorsi_cuda every step (morsi_shard_*)."""
    from imscript_b200 import shard
    M, L, check, ct = env.M, env.L, env.check, env.ct
    element, ops, w, h, planes, seed, desc = WORKLOADS["c4"]
    e = M.parse_element(element)
    op = M.OPS.index(ops[0])
    # every rank must agree that the sharded plane exists before anyone waits for a neighbour
    job, err = None, ""
    try:
        job = shard.ShardJob(L, op, e, w, h, env.rank, env.world, env.local, env.dist, seed)
    except Exception as ex:                      # e.g. no peer access / CUDA IPC between the ranks' devices
        err = str(ex)[:300]
    if env.max_over_ranks(0.0 if job is not None else 1.0) > 0:
        if job is not None:
            job.destroy()
        return {"workload": desc, "n_gpus": env.world, "scaling": "strong",
                "error": err or "another rank could not create or connect its shard"}
    sync = job.sync
    ms_per_step, launches, clocks = env.timed(job.step, job.stream, steps, warmup, sync)
    halo = int(env.sum_over_ranks(job.halo_bytes()))
    # the exchange on its own (push + wait, no kernels in between)
    up, down = M.halo_rows(op, e)
    ex_ms = None
    if env.world > 1:
        def ex():
            check(L.morsi_shard_exchange(job.s, 0, up, down))
        ex_ms, _, _ = env.timed(ex, job.stream, max(steps, 20), 3, sync)
    samples = w * h
    rec = {"workload": desc, "n_gpus": env.world, "scaling": "strong", "ms_per_step": ms_per_step,
           "value": samples / (ms_per_step * 1e-3) / 1e6, "unit": "Mpixel/s", "steps": steps, "warmup": warmup,
           "rows_per_rank": job.plan.rows_own, "halo_rows": [up, down],
           "l2": "every rank reads and writes %.1f GB per step: far beyond the 126 MB L2" % (2 * job.plan.rows_own * w * 4 / 1e9),
           "halo_bytes_per_step_all_ranks": halo, "exchange_us": None if ex_ms is None else ex_ms * 1e3,
           "exchange": "libmorsi_cuda: remote stores into the neighbours' halo rows over NVLink (CUDA IPC), "
                       "device-side credit/ready flags, interior rows overlap the transfer" if env.world > 1
                       else "single rank: no neighbours",
           "gpu_launches": launches, "clocks": clocks}
    peak, _ = peaks()
    rec["roofline_frac"] = 8.0 * samples / env.world / (ms_per_step * 1e-3) / 1e9 / peak
    if with_e2e:
        nbytes = job.host_buffers()
        e2e_steps = max(1, min(steps, 2))
        sec = env.wall(job.e2e_step, e2e_steps)
        ceil_gbs = pcie_ceiling(env, job.h_x, job.h_y, nbytes)
        rec["e2e"] = {"value": samples / sec / 1e6, "unit": "Mpixel/s",
                      "h2d_bytes_per_step": nbytes * env.world, "d2h_bytes_per_step": nbytes * env.world,
                      "steps": e2e_steps,
                      "api": "morsi_shard_apply_host (pinned host band per rank: boundary rows up and pushed first, "
                             "then upload / kernels / download pipelined in row chunks)",
                      "ceiling": {"value": ceil_gbs / 4 * 1e3, "unit": "Mpixel/s",
                                  "pcie_gbs_per_direction_all_ranks": ceil_gbs}}
        rec["e2e"]["frac_of_ceiling"] = rec["e2e"]["value"] / rec["e2e"]["ceiling"]["value"]
    job.destroy()
    return rec


def sass_counts():
    """min/max instructions per sample of the dominant kernels, from the opcode histograms
    tools/sass_histogram.py extracts from the built library (profiles/r2_sass_histogram.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_sass_histogram.json")))
    except Exception:
        return {}


def main_ours(args):
    env = Env(args)
    name = args.workload
    element, ops, w, h, planes, seed, desc = WORKLOADS[name]
    sharded_rec = None
    if name == "c4":
        sharded_rec = measure_sharded(env, args.steps, args.warmup)
        if "error" in sharded_rec:
            raise SystemExit("bench.py: the sharded workload could not be set up: " + sharded_rec["error"])
        main = {"ms_per_step": sharded_rec["ms_per_step"], "launches": sharded_rec["gpu_launches"],
                "clocks": sharded_rec["clocks"], "e2e": sharded_rec.get("e2e"), "samples_total": w * h * len(ops),
                "scaling": "strong"}
    else:
        main = measure_planes(env, name)
        if not args.no_sharded:
            # the north-star scaling curve: C4 strong scaling at this N, in every bench line
            sharded_rec = measure_sharded(env, min(args.steps, 10), 3)
    ms_per_step = main["ms_per_step"]
    value = main["samples_total"] / (ms_per_step * 1e-3) / 1e6

    if env.rank != 0:
        if env.dist is not None:
            env.dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family ---------------------------
    peak, peak_src = peaks()
    per_gpu_samples_per_op = main["samples_total"] / env.world / len(ops)
    op_ms = ms_per_step / len(ops)
    achieved = 8.0 * per_gpu_samples_per_op / (op_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_sample": 8,
                "note": "per-operation duration = step time / ops per step (CUDA events on the launching stream)"}
    # The disk and median kernels are bound by the ALU pipe, not by HBM: min/max
    # instructions per sample -- counted from the SASS of the kernel that runs, see sass_counts() --
    # against the measured FMNMX/FMNMX3 rate of 64 lanes per clock per SM.
    sc = sass_counts().get(name)
    if sc:
        sm_hz = (main["clocks"].get("sm_mhz") or 1965.0) * 1e6
        alu_peak = 64.0 * 148 * sm_hz
        alu_ach = sc["minmax_per_sample"] * per_gpu_samples_per_op / (op_ms * 1e-3)
        roofline["alu_pipe"] = {"achieved": alu_ach / 1e12, "peak": alu_peak / 1e12, "unit": "T min/max lane-instr/s",
                                "frac": alu_ach / alu_peak, "minmax_instr_per_sample": sc["minmax_per_sample"],
                                "alu_pipe_instr_per_sample": sc.get("alu_pipe_per_sample"),
                                "kernel": sc.get("kernel"),
                                "source": "profiles/r2_sass_histogram.json (tools/sass_histogram.py over the shipped .so)",
                                "peak_source": "scratch/ubench_alu.cu: FMNMX3 64 lanes/clk/SM x 148 SMs x SM clock"}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(name)
            roofline["traffic_source"] = "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full"
        except Exception:
            pass

    # ---- CPU baseline: the reference on ONE host thread, bounded sample ----
    cpu = None
    if not args.no_cpu:
        n, dt, kind, sample = run_reference_once(name, 1, big=True)
        cpu = {"value": n / dt / 1e6, "unit": "Mpixel/s", "cores": 1, "kind": kind,
               "sample": sample + f", {dt:.1f} s"}

    line = {
        "metric": "morsi Mpixel/s", "value": value, "unit": "Mpixel/s", "n_gpus": env.world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": main["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(name, env.world),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": main["e2e"], "gpu_launches": main["launches"],
        "clocks": main["clocks"],
    }
    if sharded_rec is not None and name != "c4":
        line["sharded"] = sharded_rec
    print(json.dumps(line))
    if env.dist is not None:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sharded", action="store_true", help="skip the C4 strong-scaling sub-record")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
