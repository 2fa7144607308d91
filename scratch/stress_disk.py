"""Determinism / race stress for the fused disk kernels: every shape, both CTA
sizes, repeated runs must be bit-identical to the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import imscript_b200 as M
from oracle import oracle
o = oracle()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
DISKS = ["disk2.5", "disk3", "disk3.5", "disk4", "disk4.2", "disk5", "disk5.1", "disk6", "disk7", "disk8",
         "disk9", "disk10", "disk12", "disk15"]
total = 0
for warps in ("2", "4"):
    os.environ["MORSI_DISK_W"] = warps
    for w in ((704, 600) if warps == "2" else (1100,)):
        h = 420
        x = np.stack([M.synth_host(w, h, plane=p, seed=33, dist=0) for p in range(2)])
        for name in DISKS:
            e = o.element(name)
            for op in ("opening", "bothat"):
                want = o.apply(op, e, x)
                nbad = 0
                for rep in range(reps):
                    got = M.apply(op, e, x)
                    nbad += int((got.view(np.uint32) != want.view(np.uint32)).any())
                total += nbad
                if nbad:
                    print(f"W={warps} w={w} {name} {op}: {nbad}/{reps} wrong")
print("stress done, wrong runs:", total)
