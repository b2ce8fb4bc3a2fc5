"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/dsurftomo_b200.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from dsurftomo_b200 import _lib, api, inputs


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "dsurftomo_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b(?:int|void|int64_t|double|float|const char \*)\s*\*?\s*([A-Za-z_][A-Za-z_0-9]*)\s*\(", txt)
    return sorted(set(n for n in names if n.startswith(("dsurf_", "__lsmr")) or n.endswith("_")))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.EXPORTED_SYMBOLS) <= set(syms) | {"dsurf_fatal_"}


def test_build_info_carries_the_hash_of_the_sources_in_the_tree():
    """A stale libdsurf_b200.so (built from other sources than the tree's) is detectable: the Makefile bakes the sha1
    of every library source into dsurf_build_info()."""
    import glob
    import hashlib

    csrc = os.path.join(ROOT, "dsurftomo_b200", "csrc")
    files = sorted(glob.glob(os.path.join(csrc, "*.cu")) + glob.glob(os.path.join(csrc, "*.cuh")) +
                   glob.glob(os.path.join(csrc, "*.cpp")), key=os.path.basename)
    files = [f for f in files if os.path.basename(f) != "build_info.h"]
    # GNU make's $(sort ...) orders the words as written in the Makefile: relative names, then ../../include, Makefile
    names = sorted([os.path.basename(f) for f in files] + ["../../include/dsurftomo_b200.h", "Makefile"])
    h = hashlib.sha1()
    for n in names:
        h.update(open(os.path.join(csrc, n), "rb").read())
    info = _lib.lib().dsurf_build_info().decode()
    assert info.endswith("src " + h.hexdigest()[:16]), (info, h.hexdigest()[:16])


def test_gfortran_mangled_dropins_present():
    L = _lib.lib()
    for s in ("calsurfg_", "depthkernel_", "caldespersion_", "surfdisp96_", "aprod_", "__lsmrmodule_MOD_lsmr"):
        assert hasattr(L, s)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    pb = inputs.synthetic_problem(8, 1, 2, ("Rc",), nrecv=1)
    with pytest.raises(api.DsurfError) as e:
        api.CalSurfG(pb)
    assert e.value.code == 3  # DSURF_ERR_NO_CUDA
    with pytest.raises(api.DsurfError):
        api.LSMR(2, 2, 5, 2, np.array([2, 1, 2, 1, 2], np.int32), np.ones(2, np.float32), np.ones(2, np.float32),
                 0.0, 1e-6, 1e-6, 100.0, 10, 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dsurftomo_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)


def test_synthetic_generator_shapes():
    pb = inputs.synthetic_problem(12, 3, 6, ("Rc", "Rg", "Lc", "Lg"), nrecv=5)
    assert pb.kmax == 12 and pb.ngathers == 72 and pb.dall == 72 * 5
    assert pb.nsweeps == 3 * 6 * (1 + 2 + 1 + 2)
    assert pb.scxf.shape == (12, 6) and pb.rcxf.shape == (12, 6, 6)
    # periods are 1-based indices inside each type; blocks are ordered Rc, Rg, Lc, Lg
    assert np.array_equal(pb.wavetype[:, 0], [2] * 6 + [1] * 6)
    assert np.array_equal(pb.igrt[:, 0], [0] * 3 + [1] * 3 + [0] * 3 + [1] * 3)


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference (the CPU restatement; runs without a GPU): exactly one JSON line on stdout with
    the contract's keys."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "0",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]
