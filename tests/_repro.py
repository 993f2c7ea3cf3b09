import sys; sys.path.insert(0,'.')
import numpy as np, torch, ctypes
from tests.golden_util import Golden
from tests.test_gpu_parity import make_engine
from vimure_b200 import _capi
g=Golden(sys.argv[1] if len(sys.argv)>1 else "f1_over")
eng,P=make_engine(g)
torch.cuda.synchronize()
print("init ok")
for it in range(2):
    for ph in ("gamma","phi","rho","finish"):
        eng.phase(ph, 1)
        torch.cuda.synchronize()
        print(it, ph, "ok")
    print(it, eng.elbo(), g.z["it_elbo"][it])
