"""CPU replay of the lane-per-sweep eikonal march (dsurftomo_b200/csrc/eik_lps.cuh, the code the
device kernel k_march_lps instantiates) against the oracle's Fmm::travel: bit-identical travel times
on every node, for uniform (tie-heavy), smooth and blocky velocity fields up to the 1025 x 1025 grid
of BASELINE configs[2], with the heap's shared-memory/global split placed at three different levels."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("klg", [7, 3, 2])
def test_lps_march_host_replay_is_bit_exact(tmp_path, klg):
    import oracle_lib as O

    O.lib()  # builds oracle/liboracle.so if needed
    exe = tmp_path / f"lps_check_{klg}"
    src = os.path.join(ROOT, "tests", "host", "lps_host_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", f"-DLPS_HOST_KLG={klg}", "-o", str(exe), src,
                    "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"],
                   check=True)
    r = subprocess.run([str(exe), "1" if klg == 7 else "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "LPS HOST CHECK OK" in r.stdout
    assert "mismatches=0" in r.stdout and "mismatches=1" not in r.stdout
