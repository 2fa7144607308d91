"""Opcode histograms of the dominant kernel of every bench workload, extracted from the SHIPPED
library with cuobjdump, and the min/max instructions per sample bench.py's ALU-pipe roofline uses.

    python tools/sass_histogram.py            -> profiles/r2_sass_histogram.json

k_disk: the march is unrolled over PERIOD steps, each step = 2 rows x 4 columns = 8 samples per
thread; both stage roles of a fused two-stage kernel run the same code, so the static count of
the unrolled period / (8 PERIOD) is the cost per sample AND stage.  k_median_quad: one thread
= 2 x 2 outputs.  k_small_1: one thread = 4 columns, the row loop is not unrolled.  The
counts include the few hundred instructions outside the loops (prologue, border paths)."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "imscript_b200", "lib", "libmorsi_cuda.so")
ALU = {"FMNMX3", "FMNMX", "ISETP", "VIADD", "VIMNMX3", "IADD3", "SHF", "LOP3", "LEA", "FSEL", "SEL", "PLOP3",
       "VIMNMX", "IABS", "SGXT", "VOTE", "PRMT", "FSETP", "BMSK", "FLO", "POPC"}
# workload -> (mangled-name regex, samples the counted code handles per thread, stages per operation, note)
KERNELS = {
    "c2": (r"_Z6k_diskI5ShapeILi8EELi4ELi2ELb0ELb1ELb0EE", 8 * 7, 2, "k_disk<disk7, W=2, two-stage>: PERIOD 7 steps x 8 samples"),
    "c4": (r"_Z6k_diskI5ShapeILi13EELi4ELi4ELb0ELb1ELb1EE", 8 * 15, 2, "k_disk<disk15, W=4, two-stage + x>: PERIOD 15 steps x 8 samples"),
    "c3": (r"_Z13k_median_quadILi5EE", 4, 1, "k_median_quad<disk5>: 2 x 2 outputs per thread"),
}


def histogram(pattern):
    names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn = next(m.group(1) for m in re.finditer(r"Function : (\S+)", names) if re.search(pattern, m.group(1)))
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, LIB], capture_output=True, text=True).stdout
    c = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P[0-9T]+ )?([A-Z0-9_]+)", line)
        if m:
            c[m.group(2)] += 1
    return fn, c


NCU_SOURCE = {  # ncu --import-source exports (scratch/gpu_ncu_one.sh): dynamic, per-thread instruction counts
    "c2": ("kdisk_c2_source.csv", 3 * 4096 * 4096), "c3": ("kmedian_c3_source.csv", 8192 * 8192),
    "c4": ("kdisk_c4_source.csv", 40000 * 40000),
}


def dynamic_counts(wl):
    """executed min/max thread-instructions per sample from the ncu source page of the kernel, if one was brought back"""
    import csv
    fn, samples = NCU_SOURCE[wl]
    path = os.path.join(ROOT, "gpurun_out", fn)
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    ia, it = hdr.index("Source"), hdr.index("Thread Instructions Executed")
    mm = tot = 0
    for r in data:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ia].strip())
        n = int(r[it]) if r[it].isdigit() else 0
        tot += n
        if m and m.group(2) in ("FMNMX", "FMNMX3"):
            mm += n
    return {"minmax_per_sample": mm / samples, "instructions_per_sample": tot / samples, "kernel": rows[0][1],
            "source": "gpurun_out/" + fn + " (ncu --set full --import-source on, Thread Instructions Executed)"}


def main():
    out = {}
    for wl, (pat, samples, stages, note) in KERNELS.items():
        fn, c = histogram(pat)
        mm = c["FMNMX"] + c["FMNMX3"]
        alu = sum(v for k, v in c.items() if k in ALU)
        out[wl] = {"kernel": fn, "how": note, "instructions": sum(c.values()),
                   "minmax_per_sample": stages * mm / samples, "alu_pipe_per_sample": stages * alu / samples,
                   "tma_instructions": c["UTMALDG"], "mbarrier_instructions": c["SYNCS"],
                   "histogram": dict(c.most_common(24))}
        d = dynamic_counts(wl)
        if d:
            out[wl]["dynamic"] = d
            out[wl]["static_minmax_per_sample"] = out[wl]["minmax_per_sample"]
            out[wl]["minmax_per_sample"] = d["minmax_per_sample"]      # executed counts (incl. halo / warm-up rows) win
        print(wl, fn[:60], "min/max per sample %.1f, ALU-pipe per sample %.1f, UTMALDG %d, SYNCS %d"
              % (out[wl]["minmax_per_sample"], out[wl]["alu_pipe_per_sample"], c["UTMALDG"], c["SYNCS"]))
    total = collections.Counter()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    for line in sass.splitlines():
        m = re.match(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P[0-9T]+ )?([A-Z0-9_]+)", line)
        if m:
            total[m.group(2)] += 1
    out["library"] = {"kernels": len(re.findall(r"Function : ", sass)),
                      "UTMALDG": total["UTMALDG"], "SYNCS": total["SYNCS"], "FMNMX3": total["FMNMX3"], "FMNMX": total["FMNMX"],
                      "note": "UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier operations"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_sass_histogram.json"), "w"), indent=1)
    print("library:", out["library"])


if __name__ == "__main__":
    sys.exit(main())
