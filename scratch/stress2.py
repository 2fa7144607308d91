import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import imscript_b200 as M
from oracle import oracle
o = oracle()
reps = 60
os.environ["MORSI_DISK_W"] = "2"
for w in (468, 600, 696, 704, 928):
    h = 420
    x = np.stack([M.synth_host(w, h, plane=p, seed=33, dist=0) for p in range(2)])
    for name in ("disk12", "disk10", "disk9"):
        e = o.element(name)
        for op in ("opening",):
            ref = M.apply(op, e, x)
            nbad = 0; cols = []
            for rep in range(reps):
                got = M.apply(op, e, x)
                bad = got.view(np.uint32) != ref.view(np.uint32)
                if bad.any():
                    idx = np.argwhere(bad); nbad += 1
                    cols.append((int(idx[:,2].min()), int(idx[:,2].max()), int(idx[:,1].min()), int(idx[:,1].max())))
            print(f"w={w} {name} {op}: {nbad}/{reps} differ", cols[:6])
