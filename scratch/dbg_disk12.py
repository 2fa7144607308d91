import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import imscript_b200 as M
from oracle import oracle
o = oracle()
name, op = (sys.argv[1:3] + ["disk12", "bothat"])[:2] if len(sys.argv) > 2 else ("disk12", "bothat")
for warps in ("2", "4"):
    os.environ["MORSI_DISK_W"] = warps
    h, w = 420, 1100 if warps == "4" else 600
    x = np.stack([M.synth_host(w, h, plane=p, seed=33, dist=2 if p == 1 else 0) for p in range(2)])
    x[0, 100:180, 200:330] = np.nan
    x[x == 0] = 0.0
    e = o.element(name)
    want = o.apply(op, e, x)
    for rep in range(3):
        got = M.apply(op, e, x)
        bad = ~((got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want)))
        idx = np.argwhere(bad)
        if len(idx):
            print(f"W={warps} rep {rep}: {len(idx)} bad; planes {np.unique(idx[:,0])} rows {idx[:,1].min()}..{idx[:,1].max()} cols {idx[:,2].min()}..{idx[:,2].max()}")
            rows, cnt = np.unique(idx[:, 1], return_counts=True)
            print("   rows:", list(zip(rows[:12], cnt[:12])))
            i = tuple(idx[0]); print("   first", i, got[i], want[i], "x=", x[i])
            cl = o.apply("closing", e, x)
            print("   closing there:", cl[i], " got+x =", got[i] + x[i])
        else:
            print(f"W={warps} rep {rep}: ok")
