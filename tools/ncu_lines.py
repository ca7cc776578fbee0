"""Aggregate `ncu --page source --print-source cuda,sass --csv` stall samples per CUDA source line."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, hdr, agg, tot = None, None, collections.OrderedDict(), 0
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if r[0] == 'Function Name' or hdr is None:
        continue
    if r[2] == '-' and r[0].isdigit():
        samples = int(r[4]) if r[4].isdigit() else 0
        inst = int(r[7]) if r[7].isdigit() else 0
        stalls = {hdr[i]: int(r[i]) for i in range(len(hdr)) if hdr[i].startswith('stall_') and 'Not Issued' not in hdr[i]
                  and r[i].isdigit() and int(r[i]) > 0}
        agg[(cur_file, int(r[0]))] = (samples, inst, r[1].strip()[:80], stalls)
        tot += samples
print("total samples", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(v[3].items(), key=lambda kv: -kv[1])[:3]
    print(f"{k[0]}:{k[1]:4d} {100 * v[0] / tot:5.1f}% inst={v[1]:>11d} {v[2]}   {st}")
