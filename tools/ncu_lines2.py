"""Aggregate an ncu report's source page (cuda,sass view) by source line: instruction share, stall-sample share, top stalls.
usage: python tools/ncu_lines2.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{pat}",
                      "--launch-skip", (sys.argv[4] if len(sys.argv) > 4 else "0"), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur, agg = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "" or r[2] != "-": continue
    g = lambda name: float(r[hdr.index(name)] or 0) if name in hdr else 0.0
    a = agg.setdefault((cur, int(r[0])), [0.0, 0.0, r[1].strip()[:90], {}, 0.0])
    a[0] += g("Warp Stall Sampling (All Samples)"); a[1] += g("Instructions Executed"); a[4] += g("L1 Wavefronts Shared")
    for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_lg", "stall_no_inst", "stall_math", "stall_dispatch", "stall_branch_resolving", "stall_membar", "stall_not_selected"):
        a[3][k] = a[3].get(k, 0.0) + g(k)
ti, ts = sum(a[1] for a in agg.values()), sum(a[0] for a in agg.values())
tw = sum(a[4] for a in agg.values())
print(f"total warp instructions {ti:.4g}, stall samples {ts:.4g}, shared wavefronts {tw:.4g}")
tot = {}
for a in agg.values():
    for k, v in a[3].items(): tot[k] = tot.get(k, 0) + v
print("stall mix:", {k: round(100 * v / max(sum(tot.values()), 1), 1) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:7]})
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    st = sorted(a[3].items(), key=lambda x: -x[1])[:2]
    print(f"{k[0][:12]:12s}{k[1]:5d} samp {100*a[0]/ts:5.1f}% inst {100*a[1]/ti:5.1f}% wave {100*a[4]/max(tw,1):5.1f}% {st[0][0][6:]:>9s}/{st[1][0][6:]:<9s}| {a[2]}")
