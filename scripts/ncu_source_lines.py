"""Per-source-line instruction counts of one kernel from an ncu report with --import-source on.
usage: python scripts/ncu_source_lines.py <report.ncu-rep> <rounds> [top]
`rounds` = number of warp-level loop iterations to normalise by (e.g. sweeps * nodes / 32)."""
import collections, csv, io, subprocess, sys

rep, rounds = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
iInst, iThr = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
iSt = hdr.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in hdr else None
by = collections.OrderedDict()
tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].strip():
        continue  # SASS rows repeat the counts of their source line
    try:
        inst, thr = int(r[iInst] or 0), int(r[iThr] or 0)
        st = int(r[iSt] or 0) if iSt is not None else 0
    except ValueError:
        continue
    d = by.setdefault((r[0], r[1].strip()[:100]), [0, 0, 0])
    d[0] += inst; d[1] += thr; d[2] += st
    tot_i += inst; tot_s += st
print(f"total warp instructions {tot_i:.4g} = {tot_i / rounds:.1f} per round; stall samples {tot_s}")
for (ln, src), (i, t, s) in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:>5} {i / rounds:8.1f} inst/round  {t / max(i, 1):5.1f} lanes  {100 * s / max(tot_s, 1):5.1f}% stalls  {src}")
