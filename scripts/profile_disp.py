"""ncu target: the dispersion + depth-kernel stage (K1) of the cfg-3 model, Rayleigh phase, 16 periods.
usage: python scripts/profile_disp.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsurftomo_b200 import api, inputs

pb = inputs.config(3)
t0 = time.perf_counter()
pv, svs, svp, srho = api.depthkernel(pb.nx, pb.ny, pb.nz, pb.vsf, 2, 0, pb.kmaxRc, pb.tRc, pb.depz, pb.minthk)
print("depthkernel Rc", pb.nx * pb.ny, "columns", pb.kmaxRc, "periods", time.perf_counter() - t0, "s; pv range", pv.min(), pv.max())
