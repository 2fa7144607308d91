import sys; sys.path.insert(0,'.')
import numpy as np, imscript_b200 as M
from oracle import oracle
o=oracle()
x=M.synth_host(64,64,seed=16,dist=2)
for el in ["disk2.5","disk7"]:
    e=o.element(el)
    for op in ["opening","closing","oscillation","gradient","laplacian"]:
        want=o.apply(op,e,x)
        for path in (0,1,2):
            M.lib().morsi_cuda_set_path(path)
            got=M.apply(op,e,x)
            nan=np.isnan(want)
            bad=(got.view(np.uint32)!=want.view(np.uint32))&~(nan&np.isnan(got))
            print(el,op,"path",path,"bad",int(bad.sum()), np.argwhere(bad)[:3].tolist())
M.lib().morsi_cuda_set_path(0)
