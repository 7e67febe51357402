"""Summarise gpurun_out/*.ncu-rep and launch lists into profiles/ (tracked).  Usage:
    python scripts/ncu_summary.py <tag> <launches.csv> <rep1.ncu-rep> [<rep2.ncu-rep> ...]
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    tag, launches = sys.argv[1], sys.argv[2]
    reps = sys.argv[3:]
    lines = [f"# ncu summary — {tag}", ""]
    summary = {}
    if os.path.exists(launches):
        rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
        agg = defaultdict(list)
        for r in rows:
            name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            agg[name].append(float(r[-1]))
        tot = sum(sum(v) for v in agg.values())
        lines += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
                  "| kernel | launches | mean µs | share of listed time |", "|---|---|---|---|"]
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"| `{k}` | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/tot:.3f} |")
            summary[f"share_{k}"] = sum(v) / tot
        lines.append("")
    for rep in reps:
        hdr, units, data = raw(rep)
        lines += [f"## `{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`)", ""]
        for d in data:
            name = d[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            lines += [f"### {name}", "", "| metric | value | unit |", "|---|---|---|"]
            vals = {}
            for k in WANT:
                if k in hdr:
                    i = hdr.index(k)
                    lines.append(f"| {k} | {d[i]} | {units[i]} |")
                    vals[k] = (d[i], units[i])
            lines.append("")

            def tobytes(key):
                v, u = vals.get(key, ("0", "byte"))
                f = float(v.replace(",", ""))
                return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
            short = "syrk_i8" if "syrk_i8" in name else "pack" if "pack_dosage" in name else name.split("(")[0]
            wl = tag.split("_")[-1]  # r1_c2 -> c2
            summary[f"{short}_dram_bytes_per_launch_{wl}"] = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
            summary[f"{short}_tensor_pipe_pct_{wl}"] = vals.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", ("", ""))[0]
            summary[f"{short}_dram_pct_{wl}"] = vals.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ("", ""))[0]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"ncu_{tag}.md"), "w").write("\n".join(lines) + "\n")
    sp = os.path.join(ROOT, "profiles", "ncu_summary.json")
    old = json.load(open(sp)) if os.path.exists(sp) else {}
    old.update(summary)
    old["source"] = f"profiles/ncu_{tag}.md"
    json.dump(old, open(sp, "w"), indent=1)
    print("\n".join(lines[:80]))


if __name__ == "__main__":
    main()
