"""the Fortran-free main program (dsurftomo_b200/bin/dsurftomo_b200,
source dsurftomo_b200/csrc/driver.cpp) -- readers, writers and the outer loop of src/main.f90.

CPU tests: the C++ readers against the Python mirror of main.f90:134-321 (inputs.read_problem),
the list-directed REAL*4 formatter against known gfortran output, loud failure without a GPU.
GPU test: two Taipei outer iterations through the binary against the oracle chain
(oracle CalSurfG -> host glue -> oracle LSMR -> model update)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from conftest import ROOT, has_gpu
from dsurftomo_b200 import hostglue, inputs

BIN = os.path.join(ROOT, "dsurftomo_b200", "bin", "dsurftomo_b200")
TAIPEI_IN = os.path.join(ROOT, "tests", "golden", "taipei", "DSurfTomo.in")


def _need_bin():
    if not os.path.exists(BIN):
        subprocess.run(["make", "-C", os.path.join(ROOT, "dsurftomo_b200", "csrc"), "all"], check=True,
                       capture_output=True)
    assert os.path.exists(BIN)


def test_driver_readers_match_python_mirror(tmp_path):
    """DSurfTomo.in / data file / MOD parsed by the C++ driver == inputs.read_problem, bit for bit
    (REAL*4 colatitude/longitude conversion, delsph distances, obst = dist / velocity)."""
    _need_bin()
    dump = tmp_path / "parsed.bin"
    r = subprocess.run([BIN, TAIPEI_IN, "--outdir", str(tmp_path), "--parse-only", str(dump), "--quiet"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pb = inputs.read_problem(TAIPEI_IN)
    raw = open(dump, "rb").read()
    hdr = struct.unpack("12i", raw[:48])
    fh = np.frombuffer(raw[48:96], np.float32)
    (maxnar,) = struct.unpack("q", raw[96:104])
    assert hdr == (pb.nx, pb.ny, pb.nz, pb.nsrc, pb.kmax, pb.dall, pb.kmaxRc, pb.kmaxRg, pb.kmaxLc, pb.kmaxLg,
                   pb.maxiter, pb.ifsyn)
    want = np.array([pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pb.weight, pb.damp, pb.minthk, pb.minvel, pb.maxvel,
                     pb.spfra, pb.noiselevel, pb.threshold], np.float32)
    assert np.array_equal(fh, want)
    assert maxnar == pb.maxnar()
    off = 104

    def take(n, dt):
        nonlocal off
        a = np.frombuffer(raw[off:off + n * np.dtype(dt).itemsize], dt)
        off += n * np.dtype(dt).itemsize
        return a

    for t in (pb.tRc, pb.tRg, pb.tLc, pb.tLg):
        assert np.array_equal(take(len(t), np.float64), t)
    for a in (pb.scxf, pb.sczf, pb.rcxf, pb.rczf):
        assert np.array_equal(take(a.size, np.float32), a.ravel())
    for a in (pb.periods, pb.wavetype, pb.igrt, pb.nrc1, pb.nsrc1):
        assert np.array_equal(take(a.size, np.int32), a.ravel())
    for a in (pb.obst, pb.dist, pb.depz, pb.vsf):
        assert np.array_equal(take(a.size, np.float32), a.ravel())
    assert off == len(raw)
    log = open(tmp_path / "DSurfTomo.in.log").read().splitlines()
    assert log[1].strip() == "S U R F  T O M O" and log[5] == "  25.20000 121.35000"  # '(2f10.5)'
    assert log[9] == "   18   18    9"                                               # '(3i5)'
    assert log[11] == "".join("%7.2f" % t for t in pb.tRc)                           # '(50f7.2)'


def test_driver_list_directed_real_format():
    """write(88,*) of REAL*4 (residualFirst/Last.dat, DWS line): gfortran prints G16.9E2-style items
    separated by one blank; known outputs of `print *, x`."""
    _need_bin()
    vals = ["1", "0.1", "123.456", "1e-5", "1e10", "0", "-2.5", "31.5"]
    r = subprocess.run([BIN, "--ld-real", *vals], capture_output=True, text=True)
    got = [l[1:-1] for l in r.stdout.splitlines()]
    assert got == ["   1.00000000    ", "  0.100000001    ", "   123.456001    ", "   9.99999975E-06",
                   "   1.00000000E+10", "   0.00000000    ", "  -2.50000000    ", "   31.5000000    "]
    assert all(float(g) == np.float32(v) for g, v in zip(got, vals))  # nine digits round-trip REAL*4


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_driver_fails_loudly_without_gpu(tmp_path):
    _need_bin()
    r = subprocess.run([BIN, TAIPEI_IN, "--outdir", str(tmp_path), "--maxiter", "1", "--quiet"],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


def _read_model_file(path, pb):
    a = np.loadtxt(path)
    assert a.shape == ((pb.nx - 2) * (pb.ny - 2) * (pb.nz - 1), 4)
    return a


@pytest.mark.gpu
def test_driver_two_outer_iterations_match_oracle_chain(taipei, tmp_path):
    _need_bin()
    pb = taipei
    r = subprocess.run([BIN, TAIPEI_IN, "--outdir", str(tmp_path), "--maxiter", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "Program finishes successfully" in r.stdout and " 2th iteration..." in r.stdout
    # oracle chain
    vs = pb.vsf.copy()
    models, first = [], None
    for it in range(2):
        ref = O.calsurfg(pb, vels=vs, nthreads=8, mode=1)
        s = hostglue.host_glue(pb, ref["dsurf"], ref["row"], ref["col"], ref["rw"])
        if it == 0:
            first = (ref["dsurf"].copy(), s["datweight"].copy())
        L = O.lsmr(s["m"], s["n"], O.pack_iw(s["rows"], s["cols"]), s["vals"], s["cbst"], pb.damp)
        vs, _ = hostglue.model_update(pb, vs, L["x"])
        models.append(vs.copy())
    lon = pb.gozd + np.arange(pb.ny - 2) * pb.dvzd
    lat = pb.goxd - np.arange(pb.nx - 2) * pb.dvxd
    for it, name in ((0, "DSurfTomo.inMeasure.dat.iter001"), (1, "DSurfTomo.inMeasure.dat.iter002"),
                     (1, "DSurfTomo.inMeasure.dat")):
        a = _read_model_file(tmp_path / name, pb)
        want = models[it][: pb.nz - 1, 1:-1, 1:-1]                       # [k][j][i], i fastest in the file
        got = a[:, 3].reshape(pb.nz - 1, pb.ny - 2, pb.nx - 2)
        # '(f10.5)' rounding (5e-6 absolute) + parity: 1e-5 relative after one outer iteration; the second
        # iteration starts from models that already differ by 1e-5, LSMR amplifies that to a few 1e-5
        rel = 1e-5 if it == 0 else 5e-5
        err = np.abs(got - want) - 5.1e-6
        assert np.all(err <= rel * want), (name, float((err / want).max()))
        assert np.allclose(a[: pb.nx - 2, 1], lat, atol=6e-6) and np.allclose(a[:: pb.nx - 2, 0][: pb.ny - 2], lon, atol=6e-6)
        assert np.allclose(a[:: (pb.nx - 2) * (pb.ny - 2), 2], pb.depz[: pb.nz - 1], atol=6e-6)
    # residualFirst.dat: dist, dsyn, obst, dsyn*w, obst*w, w  (main.f90:396-403)
    res = np.loadtxt(tmp_path / "residualFirst.dat")
    assert res.shape == (pb.dall, 6)
    assert np.array_equal(res[:, 0].astype(np.float32), pb.dist) and np.array_equal(res[:, 2].astype(np.float32), pb.obst)
    assert np.abs(res[:, 1] / first[0] - 1).max() <= 1e-5
    assert (res[:, 5].astype(np.float32) != first[1]).sum() <= 2  # outlier cut on 1e-5-different residuals
    assert os.path.exists(tmp_path / "residualLast.dat")
    log = open(tmp_path / "DSurfTomo.in.log").read()
    assert log.count("Maximum and Average DWS values:") == 2 and log.count("th iteration...") == 2
    assert "min and max velocity variation" in log and "Program finishes successfully" in log


@pytest.mark.gpu
def test_driver_checkerboard_run(taipei, tmp_path):
    """ifsyn = 1 (main.f90:323-343, 553-573): MOD.true is read, `synthetic` replaces the observed times by forward
    times through it (noiselevel 0 -> deterministic), one outer iteration runs on them, and Vs_model.real /
    <input>Syn.dat are written.  Checked against the oracle's `synthetic` and the files' contents."""
    import shutil

    _need_bin()
    pb = taipei
    src = os.path.dirname(TAIPEI_IN)
    for f in ("surfdataTB.dat", "MOD"):
        shutil.copy(os.path.join(src, f), tmp_path / f)
    lines = open(TAIPEI_IN).read().splitlines()
    lines[18] = "1                                c: synthetic flag(0:real data,1:synthetic)"
    lines[19] = "0.0                              c: noiselevel"
    (tmp_path / "DSurfTomo.in").write_text("\n".join(lines) + "\n")
    ii = np.arange(pb.nx)[None, None, :]
    jj = np.arange(pb.ny)[None, :, None]
    vtrue = np.clip(pb.vsf + 0.1 * np.sin(0.7 * ii) * np.sin(0.7 * jj), pb.minvel, pb.maxvel).astype(np.float32)
    with open(tmp_path / "MOD.true", "w") as fh:  # no depth line (main.f90:330-336)
        for k in range(pb.nz):
            for j in range(pb.ny):
                fh.write(" ".join("%.5f" % v for v in vtrue[k, j]) + "\n")
    vtrue = np.array([[["%.5f" % v for v in row] for row in pl] for pl in vtrue], dtype=np.float64).astype(np.float32)
    r = subprocess.run([BIN, str(tmp_path / "DSurfTomo.in"), "--maxiter", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "Synthetic Test Begin" in r.stdout and "Vs_model.real" in r.stdout
    ref = O.synthetic(pb, vels=vtrue, nthreads=8)
    assert ref["err"] == 0
    res = np.loadtxt(tmp_path / "residualFirst.dat")
    assert np.abs(res[:, 2] / ref["obst"] - 1).max() <= 1e-5          # obst column = synthetic forward times
    real = _read_model_file(tmp_path / "Vs_model.real", pb)
    want = vtrue[: pb.nz - 1, 1:-1, 1:-1].ravel()
    assert np.abs(real[:, 3] - want).max() <= 5.1e-6
    syn = _read_model_file(tmp_path / "DSurfTomo.inSyn.dat", pb)
    it1 = _read_model_file(tmp_path / "DSurfTomo.inMeasure.dat.iter001", pb)
    assert np.array_equal(syn, it1) and not os.path.exists(tmp_path / "DSurfTomo.inMeasure.dat")
    for name in ("velmap2dRc.dat",):
        assert os.path.exists(tmp_path / name)
    # the inversion moved the start model towards the checkerboard
    start = pb.vsf[: pb.nz - 1, 1:-1, 1:-1].ravel()
    assert np.abs(syn[:, 3] - want).mean() < np.abs(start - want).mean()


def test_driver_reader_list_directed_edge_cases(tmp_path):
    """List-directed input forms gfortran accepts in DSurfTomo.in / the data file: comma separators, D exponents,
    period lists wrapped over several records, trailing comments, blank lines in the data file; and the reference's
    STOP on a missing input file (main.f90:130-131)."""
    _need_bin()
    (tmp_path / "DSurfTomo.in").write_text(
        "c\nc\nc\n"
        "data.dat   c: data file\n"
        "5, 6, 3    c: nx ny nz\n"
        "25.2d0 , 121.35   c: origin\n"
        "1.5D-2 0.017\n"
        "4\n"
        "4.0,1.0\n"
        "3\n"
        "0.5 2.8\n"
        "2\n"
        "0.2\n"
        "3      c: kmaxRc\n"
        "0.5 1.0\n"
        "2.0    c: wrapped period list\n"
        "0\n0\n0\n"
        "0\n0.02\n3.0\n")
    (tmp_path / "data.dat").write_text(
        "# 25.18 121.37 1 2 0\n"
        "25.17, 121.39, 1.25\n"
        "\n"
        "25.175 121.40 1.3d0\n"
        "# 25.17 121.39 3 2 0\n"
        "25.18 121.37 1.5\n")
    (tmp_path / "MOD").write_text("0.0 0.5 1.0\n" + ("1.0 1.1 1.2 1.3 1.4\n" * 6 + "2.0 2.1 2.2 2.3 2.4\n" * 6 +
                                                    "3.0 3.1 3.2 3.3 3.4\n" * 6))
    dump = tmp_path / "p.bin"
    r = subprocess.run([BIN, str(tmp_path / "DSurfTomo.in"), "--parse-only", str(dump), "--quiet"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(dump, "rb").read()
    hdr = struct.unpack("12i", raw[:48])
    assert hdr[:10] == (5, 6, 3, 4, 3, 3, 3, 0, 0, 0)          # nx ny nz nsrc kmax dall kmaxRc..Lg
    fh = np.frombuffer(raw[48:96], np.float32)
    assert fh[0] == np.float32(25.2) and fh[2] == np.float32(0.015) and fh[4] == 4.0 and fh[5] == 1.0
    t = np.frombuffer(raw[104:104 + 24], np.float64)
    assert np.array_equal(t, [0.5, 1.0, 2.0])
    off = 104 + 24 + 4 * (2 * 4 * 3 + 2 * 4 * 4 * 3)          # periods, scxf/sczf, rcxf/rczf
    ints = np.frombuffer(raw[off:off + 4 * (4 * 4 * 3 + 3)], np.int32)
    nrc1, nsrc1 = ints[3 * 12:4 * 12].reshape(3, 4), ints[4 * 12:]
    assert nsrc1.tolist() == [1, 0, 1] and nrc1[0, 0] == 2 and nrc1[2, 0] == 1
    obst = np.frombuffer(raw[off + ints.nbytes:off + ints.nbytes + 12], np.float32)
    pb_dist = inputs.delsph(np.float32((90 - np.float32(25.18)) * inputs.PI32 / np.float32(180)),
                            np.float32(np.float32(121.37) * inputs.PI32 / np.float32(180)),
                            np.float32((90 - np.float32(25.17)) * inputs.PI32 / np.float32(180)),
                            np.float32(np.float32(121.39) * inputs.PI32 / np.float32(180)))
    assert obst[0] == np.float32(pb_dist / np.float32(1.25))
    r = subprocess.run([BIN, str(tmp_path / "missing.in")], capture_output=True, text=True)
    assert r.returncode == 2 and "unable to open the inputfile" in r.stderr
