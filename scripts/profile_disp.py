"""ncu / A-B target: the dispersion + depth-kernel stage (K1) of the cfg-3 model, Rayleigh phase, 16 periods.
usage: [DSURF_DISP_OTF=0|8|9|10|12] python scripts/profile_disp.py [reps]   (prints the best wall time of `reps` calls
and a checksum of the outputs; the variants must agree bit for bit)"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsurftomo_b200 import api, inputs

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pb = inputs.config(3)
rng = np.random.default_rng(5)
vs = (pb.vsf * (1.0 + 0.03 * rng.standard_normal(pb.vsf.shape))).astype(np.float32)  # laterally varying columns
best = 1e9
for _ in range(reps):
    t0 = time.perf_counter()
    pv, svs, svp, srho = api.depthkernel(pb.nx, pb.ny, pb.nz, vs, 2, 0, pb.kmaxRc, pb.tRc, pb.depz, pb.minthk)
    best = min(best, time.perf_counter() - t0)
h = hashlib.sha256()
for a in (pv, svs, svp, srho):
    h.update(np.ascontiguousarray(a).tobytes())
print("variant", os.environ.get("DSURF_DISP_OTF", "default"), "depthkernel Rc", pb.nx * pb.ny, "columns", pb.kmaxRc,
      "periods best", round(best, 4), "s sha256", h.hexdigest()[:16], "pv range", float(pv.min()), float(pv.max()))
