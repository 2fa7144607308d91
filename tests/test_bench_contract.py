"""bench.py's reference arm runs on the host cores alone: check here (no GPU)
that it prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("square erosion 1024x1024")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_ours_refuses_without_a_gpu():
    """no CPU fallback: without a CUDA device the product arm must fail loudly, not measure the oracle"""
    import imscript_b200 as M
    if M.device_count() > 0:
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True,
                         text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)


def test_reference_arm_never_maps_the_cuda_library():
    """the reference arm is the reference's CPU code alone: libmorsi_cuda.so must not be loaded by it"""
    code = ("import sys; sys.path.insert(0, %r); import bench; "
            "bench.run_reference_once('c1', 1); "
            "print('MAPPED' if 'libmorsi_cuda' in open('/proc/self/maps').read() else 'CLEAN')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip().endswith("CLEAN")


def test_numpy_synth_equals_the_library_synth():
    import numpy as np
    import bench
    import imscript_b200 as M
    for (w, rows, row0, plane, seed) in [(1024, 768, 1664, 3, 2), (37, 5, 0, 0, 1), (40000, 3, 39990, 0, 4)]:
        a = bench.synth_numpy(w, rows, row0, plane, seed)
        b = M.synth_host(w, rows, row0=row0, plane=plane, seed=seed)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
