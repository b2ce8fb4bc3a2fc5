# round 2, call 17: fast-iterative eikonal -- deviation table at 1025^2 (cfg-3 slice) and on the checkerboard worst case;
# ncu metrics of the three kernels at one resident wave (296 sweeps, 1025^2) and per-line instruction counts
mkdir -p gpurun_out
timeout 900 python scripts/fim_parity.py cfg3slice:1 checker > gpurun_out/r2s17_fim_parity.jsonl 2> gpurun_out/r2s17_fim_parity.err
cat gpurun_out/r2s17_fim_parity.jsonl; tail -n 5 gpurun_out/r2s17_fim_parity.err
DSURF_EIKONAL=fim timeout 600 ncu --metrics $(cat scripts/ncu_eik_metrics.txt) --clock-control none -k regex:"k_fim|k_refine" --csv --log-file gpurun_out/r2s17_fim_metrics.csv python scripts/profile_eikonal.py 131 296 1 > gpurun_out/r2s17_prof.log 2>&1
tail -n 2 gpurun_out/r2s17_prof.log
DSURF_EIKONAL=fim timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on --clock-control none -k regex:k_fim_march -c 1 -f -o gpurun_out/r2s17_fim_src python scripts/profile_eikonal.py 131 296 1 > gpurun_out/r2s17_prof2.log 2>&1
tail -n 2 gpurun_out/r2s17_prof2.log
