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
("strong").  torch is used for the process group only.
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
    Returns (samples processed, seconds, kind)."""
    import imscript_b200 as M
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
    crops = [M.synth_host(cw, ch, row0=(h - ch) // 2, plane=t, seed=seed + seed_offset) for t in range(threads)]

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


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    element, ops, w, h, planes, seed, desc = WORKLOADS[name]
    threads = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        run_reference_once(name, threads)
    total, secs = 0, 0.0
    for k in range(args.steps):
        n, dt, kind, sample = run_reference_once(name, threads, seed_offset=k)
        total += n
        secs += dt
    value = total / secs / 1e6
    print(json.dumps({
        "impl": "reference", "metric": "morsi Mpixel/s", "value": value, "unit": "Mpixel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "element": element, "ops": ops},
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": threads, "kind": kind,
                         "sample": sample + " per step, one crop per host thread"},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
def main_ours(args):
    import imscript_b200 as M
    from imscript_b200.binding import check
    L = M.lib()
    world, rank, local, dist, torch = dist_setup(args.gpus)
    if M.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libmorsi_cuda has no CPU fallback")
    check(L.morsi_cuda_init(local))
    name = args.workload
    element, ops, w, h, planes, seed, desc = WORKLOADS[name]
    e = M.parse_element(element)
    e_p = e.ctypes.data_as(M.binding._i32p)
    opi = [M.OPS.index(o) for o in ops]
    sharded = name == "c4"

    stream = None
    if sharded:
        from imscript_b200 import shard
        job = shard.BandJob(L, opi[0], e, w, h, rank, world, dist, torch, seed)
        samples_per_step_total = w * h * len(ops)
        step = job.step
        scaling = "strong"
        stream = job.stream
    else:
        n = w * h * planes
        d_x = M.DeviceBuffer(n * 4)
        d_y = M.DeviceBuffer(n * 4)
        for p in range(planes):
            check(L.morsi_cuda_synth(d_x.ptr + p * w * h * 4, w, h, 0, p + rank * planes, seed, 0, None))
        check(L.morsi_cuda_sync(None))
        samples_per_step_total = n * len(ops) * world
        scaling = "weak"

        def step():
            for o in opi:
                check(L.morsi_cuda_apply_device(o, e_p, d_x.ptr, d_y.ptr, w, h, planes, None))

    def barrier():
        check(L.morsi_cuda_sync(stream))
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    ev = [M.binding._vp() for _ in range(2)]
    for x in ev:
        check(L.morsi_cuda_event_create(M.binding.ctypes.byref(x)))

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.load_begin()
    L.morsi_cuda_launch_count_reset()
    check(L.morsi_cuda_event_record(ev[0], stream))
    for _ in range(args.steps):
        step()
    check(L.morsi_cuda_event_record(ev[1], stream))
    check(L.morsi_cuda_sync(stream))
    launches = L.morsi_cuda_launch_count()
    ms = M.binding.ctypes.c_float()
    check(L.morsi_cuda_event_elapsed_ms(ev[0], ev[1], M.binding.ctypes.byref(ms)))
    barrier()
    elapsed_ms = ms.value
    if dist is not None:
        t = torch.tensor([elapsed_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_per_step = elapsed_ms / args.steps
    value = samples_per_step_total / (ms_per_step * 1e-3) / 1e6
    # keep the same load running (untimed) until the clock sampler has seen it;
    # under torchrun every rank runs the same number of extra steps (a step may
    # exchange halo rows with its neighbours)
    if dist is not None:
        for _ in range(min(5000, int(200.0 / max(ms_per_step, 1e-3)) + 1)):
            step()
    else:
        t_more = time.time()
        while sampler.proc and sampler.samples_under_load() < 5 and time.time() - t_more < 1.5:
            for _ in range(max(1, args.steps // 4)):
                step()
            check(L.morsi_cuda_sync(stream))
    barrier()
    clocks = sampler.stop()

    # ---- e2e: the public host-pointer call, pinned host buffers ----------
    e2e = None
    if sharded:
        # the band lives in pinned host memory: H2D of the owned rows, halo
        # exchange, kernels, D2H of the result (1 GPU: 6.4 GB each way per step)
        nbytes = job.host_buffers()
        e2e_steps = max(1, min(args.steps, 3))
        job.e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            job.e2e_step()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": samples_per_step_total * e2e_steps / dt / 1e6, "unit": "Mpixel/s",
               "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world, "steps": e2e_steps,
               "api": "pinned host band -> morsi_cuda_memcpy_h2d + halo exchange + "
                      "morsi_cuda_apply_band_device + morsi_cuda_memcpy_d2h (per rank)"}
        job.free_host_buffers()
    if not sharded:
        hx, hy = M.binding._vp(), M.binding._vp()
        nbytes = w * h * planes * 4
        check(L.morsi_cuda_host_alloc(M.binding.ctypes.byref(hx), nbytes))
        check(L.morsi_cuda_host_alloc(M.binding.ctypes.byref(hy), nbytes))
        check(L.morsi_cuda_memcpy_d2h(hx, d_x.ptr, nbytes, None))
        check(L.morsi_cuda_sync(None))
        e2e_steps = max(1, min(args.steps, 5))

        def e2e_step():
            for o in opi:
                check(L.morsi_cuda_apply(o, e_p, hx, hy, w, h, planes))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": samples_per_step_total * e2e_steps / dt / 1e6, "unit": "Mpixel/s",
               "h2d_bytes_per_step": nbytes * len(ops), "d2h_bytes_per_step": nbytes * len(ops),
               "steps": e2e_steps, "api": "morsi_cuda_apply (host pointers, pinned)"}
        L.morsi_cuda_host_free(hx)
        L.morsi_cuda_host_free(hy)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family ---------------------------
    peak, peak_src = peaks()
    per_gpu_samples_per_op = samples_per_step_total / world / len(ops)
    op_ms = ms_per_step / len(ops)
    achieved = 8.0 * per_gpu_samples_per_op / (op_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_sample": 8,
                "note": "per-operation duration = step time / ops per step (CUDA events on the launching stream)"}
    # The disk and median kernels are bound by the ALU pipe, not by HBM:
    # min/max instructions per sample (counted in the SASS of the kernel that runs)
    # against the measured FMNMX/FMNMX3 rate of 64 lanes per clock per SM.
    minmax_per_sample = {"c2": 2 * 12.8, "c4": 2 * 27.0, "c3": 333.0}.get(name)
    if minmax_per_sample:
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        alu_peak = 64.0 * 148 * sm_hz
        alu_ach = minmax_per_sample * per_gpu_samples_per_op / (op_ms * 1e-3)
        roofline["alu_pipe"] = {"achieved": alu_ach / 1e12, "peak": alu_peak / 1e12, "unit": "T min/max lane-instr/s",
                                "frac": alu_ach / alu_peak, "minmax_instr_per_sample": minmax_per_sample,
                                "peak_source": "scratch/ubench_alu.cu: FMNMX3 64 lanes/clk/SM x 148 SMs x SM clock"}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(name)
        except Exception:
            pass

    # ---- CPU baseline: the reference on ONE host thread, bounded sample ----
    cpu = None
    if not args.no_cpu:
        n, dt, kind, sample = run_reference_once(name, 1, big=True)
        cpu = {"value": n / dt / 1e6, "unit": "Mpixel/s", "cores": 1, "kind": kind,
               "sample": sample + f", {dt:.1f} s"}

    print(json.dumps({
        "metric": "morsi Mpixel/s", "value": value, "unit": "Mpixel/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": desc, "element": element, "ops": ops, "samples_per_step": samples_per_step_total,
                   "l2": "input+output per op exceed the 126 MB L2" if w * h * planes * 8 > 126e6
                   else "working set fits L2 (launch-latency-bound config)"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
