"""Independent pins of the oracle's multi-layer and second-order paths (missing #5): nothing here
shares code or formulas with the oracle / the reference.

* Rayleigh and Love phase velocities of a 5-layer stack (surfdisp96.f:767-1062 dltar4/var/dnka/normc, :704 dltar1)
  against the roots of a secular function built from the equations of motion themselves: the motion-stress
  vector is propagated through every layer with the matrix exponential of the first-order system
  (Aki & Richards, Quantitative Seismology, eqs 7.24 / 7.28), starting from the decaying eigenvectors of the
  half-space, and the free-surface tractions must vanish.
* Group velocities (surfdisp96.f:281-300) against the same numerical differentiation done in float64 on the
  independent roots.
* Every two-dimensional branch of fouds2 (CalSurfG.f90:664-738: second/second, second/first, first/second,
  first/first order) on a travel-time field that is linear in the grid coordinates: one-sided first- and
  second-order differences are exact for it, so the update must return the field's own value at the node.
"""
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg
import scipy.optimize

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# layer thickness (km), Vp, Vs (km/s), density (g/cm^3); last = half-space
STACK = np.array([[0.4, 2.1, 1.00, 2.00],
                  [0.6, 2.9, 1.45, 2.15],
                  [0.8, 3.4, 1.30, 2.25],   # a low-velocity layer
                  [1.2, 4.3, 2.30, 2.40],
                  [1.5, 5.2, 2.85, 2.55],
                  [0.0, 6.1, 3.50, 2.75]])


def _rayleigh_system(k, w, a, b, rho):
    mu, lam = rho * b * b, rho * (a * a - 2 * b * b)
    l2m = lam + 2 * mu
    zeta = 4 * mu * (lam + mu) / l2m
    return np.array([[0.0, k, 1.0 / mu, 0.0],
                     [-k * lam / l2m, 0.0, 0.0, 1.0 / l2m],
                     [k * k * zeta - w * w * rho, 0.0, 0.0, k * lam / l2m],
                     [0.0, -w * w * rho, -k, 0.0]])


def _love_system(k, w, b, rho):
    mu = rho * b * b
    return np.array([[0.0, 1.0 / mu], [k * k * mu - w * w * rho, 0.0]])


def _secular(c, T, wave):
    """Free-surface traction determinant of the solutions that decay in the half-space (real, sign-changing at roots)."""
    w = 2 * np.pi / T
    k = w / c
    thk, a, b, rho = STACK[:, 0], STACK[:, 1], STACK[:, 2], STACK[:, 3]
    A = _rayleigh_system(k, w, a[-1], b[-1], rho[-1]) if wave == "R" else _love_system(k, w, b[-1], rho[-1])
    ev, vec = np.linalg.eig(A)
    dec = np.argsort(ev.real)[: (2 if wave == "R" else 1)]  # eigenvalues with negative real part: decay with depth
    Y = vec[:, dec].real
    assert np.all(ev.real[dec] < 0) and np.allclose(vec[:, dec].imag, 0)
    for i in range(len(thk) - 2, -1, -1):  # upwards: y(z - d) = expm(-A d) y(z)
        A = _rayleigh_system(k, w, a[i], b[i], rho[i]) if wave == "R" else _love_system(k, w, b[i], rho[i])
        Y = scipy.linalg.expm(-A * thk[i]) @ Y
        Y = Y / np.abs(Y).max()  # positive rescaling keeps the sign of the determinant
    return float(np.linalg.det(Y[2:4, :])) if wave == "R" else float(Y[1, 0])


def _fundamental_root(T, wave, lo, hi):
    cs = np.linspace(lo, hi, 1500)
    f = np.array([_secular(c, T, wave) for c in cs])
    i = np.nonzero(np.sign(f[:-1]) != np.sign(f[1:]))[0][0]  # first sign change from below = fundamental mode
    return scipy.optimize.brentq(lambda c: _secular(c, T, wave), cs[i], cs[i + 1], xtol=1e-13, rtol=1e-13)


def _oracle(iwave, igr, t):
    thk, vp, vs, rho = (STACK[:, j].astype(np.float32) for j in range(4))
    c, nf = O.surfdisp96(thk, vp, vs, rho, 0, iwave, 1, igr, np.asarray(t, np.float64))
    assert nf == 0
    return c


PERIODS = [0.4, 0.7, 1.0, 1.6, 2.5, 4.0]


@pytest.mark.parametrize("wave,iwave", [("R", 2), ("L", 1)])
def test_five_layer_phase_velocity_matches_independent_propagator(wave, iwave):
    c = _oracle(iwave, 0, PERIODS)
    bmin, bmax = STACK[:, 2].min(), STACK[-1, 2]
    for T, ck in zip(PERIODS, c):
        root = _fundamental_root(T, wave, 0.7 * bmin, 0.9999 * bmax)
        # the reference refines roots to 1e-6 relative and rounds the result to REAL*4 (surfdisp96.f:292-297, 583, 608)
        assert abs(ck - root) <= 3e-6 * root, (wave, T, ck, root)


@pytest.mark.parametrize("wave,iwave", [("R", 2), ("L", 1)])
def test_five_layer_group_velocity_matches_independent_propagator(wave, iwave):
    u = _oracle(iwave, 1, PERIODS)
    bmin, bmax = STACK[:, 2].min(), STACK[-1, 2]
    h = 0.005  # surfdisp96.f:120
    for T, uk in zip(PERIODS, u):
        ta, tb = T / (1 + h), T / (1 - h)
        ca = _fundamental_root(ta, wave, 0.7 * bmin, 0.9999 * bmax)
        cb = _fundamental_root(tb, wave, 0.7 * bmin, 0.9999 * bmax)
        ug = (1 / ta - 1 / tb) / (1 / (ta * ca) - 1 / (tb * cb))  # :300, in float64 here
        # the reference forms this quotient in REAL*4 from roots known to 1e-6: differentiation over a 1 % period
        # step amplifies that to a few 1e-4
        assert abs(uk - ug) <= 6e-4 * ug, (wave, T, uk, ug)


def test_fouds2_two_dimensional_branches_are_exact_for_a_linear_field(tmp_path):
    O.lib()
    exe = tmp_path / "fouds2_pin"
    src = os.path.join(ROOT, "tests", "host", "fouds2_pin.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-o", str(exe), src,
                    "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"],
                   check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    rows = [l.split() for l in r.stdout.splitlines() if l.startswith("case")]
    assert len(rows) == 16  # 4 branches x 4 quadrants
    for row in rows:
        rel = float(row[-1])
        assert rel <= 4e-6, row  # REAL*4 evaluation of the quadratic around T ~ 100 s with steps of ~0.1 s
