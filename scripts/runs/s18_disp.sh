mkdir -p gpurun_out
for v in 0 8 10 12; do DSURF_DISP_OTF=$v timeout 200 python scripts/profile_disp.py 2 2>&1 | tail -n 1; done | tee gpurun_out/s18_disp_ab.log
( timeout 600 python -m pytest tests -m gpu -q -k "on_the_fly" ) > gpurun_out/s18_pytest.log 2>&1; tail -n 5 gpurun_out/s18_pytest.log
DSURF_DISP_OTF=10 timeout 300 python -m pytest tests -m gpu -q -k "depthkernel or surfdisp or calsurfg_small" 2>&1 | tail -n 3
