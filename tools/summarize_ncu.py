"""Turn the .ncu-rep files in gpurun_out/ into small tracked summaries under profiles/ (per round).

    python tools/summarize_ncu.py r1 gpurun_out/prof_r1.ncu-rep gpurun_out/prof_stream_r1.ncu-rep
"""
import csv, io, json, os, subprocess, sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active"]
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    summary, traffic = [], {}
    for rep in reps:
        hdr, units, rows = raw(rep)
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows:
            name = r[ix["Kernel Name"]].strip()
            d = OrderedDict(kernel=name, report=os.path.basename(rep))
            for k in KEYS:
                if k in ix and r[ix[k]] != "":
                    v, u = r[ix[k]], units[ix[k]]
                    try:
                        v = float(v.replace(",", ""))
                    except ValueError:
                        pass
                    if u in BYTES and isinstance(v, float):
                        v, u = v * BYTES[u], "byte"
                    if k == "gpu__time_duration.sum" and u in TIME and isinstance(v, float):
                        v, u = v * TIME[u], "ms"
                    d[k] = {"value": v, "unit": u}
            st = {h.split("stalled_")[1].split("_per")[0]: float(r[ix[h]]) for h in hdr
                  if "issue_stalled" in h and "per_issue_active" in h and r[ix[h]] not in ("", "n/a")}
            d["stall_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:6])
            summary.append(d)
            dram = d.get("dram__bytes_read.sum", {}).get("value", 0) + d.get("dram__bytes_write.sum", {}).get("value", 0)
            traffic.setdefault(name, []).append({"dram_bytes": dram, "ms": d["gpu__time_duration.sum"]["value"]})
    os.makedirs("profiles", exist_ok=True)
    json.dump(summary, open(f"profiles/{tag}_ncu_summary.json", "w"), indent=1)
    with open(f"profiles/{tag}_ncu_summary.md", "w") as f:
        f.write(f"# ncu --set full summaries ({tag}); one row per captured launch\n\n")
        f.write("| kernel | ms | DRAM rd+wr (MB) | regs | grid x block | smem wavefronts | fp64 pipe % | issue % | top stalls (per issue) |\n|---|---|---|---|---|---|---|---|---|\n")
        for d in summary:
            g = lambda k: d.get(k, {}).get("value", float("nan"))
            dram = (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) / 1e6
            short = d["kernel"].replace("tfin::", "").split("(")[0]
            f.write(f"| `{short}` | {g('gpu__time_duration.sum'):.3f} | {dram:.2f} | {g('launch__registers_per_thread'):.0f} | "
                    f"{g('launch__grid_size'):.0f} x {g('launch__block_size'):.0f} | {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):.3g} | "
                    f"{g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                    f"{', '.join(f'{k} {v:.2f}' for k, v in d['stall_per_issue'].items())} |\n")
    print(open(f"profiles/{tag}_ncu_summary.md").read())
    return traffic


if __name__ == "__main__":
    main()
