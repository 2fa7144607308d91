import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import imscript_b200 as M
from oracle import oracle
o = oracle()
reps = 80
os.environ["MORSI_DISK_W"] = "2"
w, h = 704, 420
x = np.stack([M.synth_host(w, h, plane=p, seed=33, dist=0) for p in range(2)])
for knobs in ("0", "1", "2", "4", "7"):
    os.environ["MORSI_DISK_KNOBS"] = knobs
    for name in ("disk12", "disk10"):
        e = o.element(name)
        want = o.apply("opening", e, x)
        nbad = 0
        for rep in range(reps):
            got = M.apply("opening", e, x)
            nbad += int((got.view(np.uint32) != want.view(np.uint32)).any())
        print(f"knobs={knobs} {name}: {nbad}/{reps} wrong")
