"""Turns ncu output brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python scripts/ncu_summary.py full gpurun_out/prof_trace_r1.ncu-rep profiles/r1_trace_full.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): gpu__time_duration.sum per kernel, --clock-control none\n\n")
        f.write("Per-launch times are cold-cache and serialised by the profiler: read the SHARES.\n\n")
        f.write("| kernel | launches | mean us | share of all profiled time |\n|---|---:|---:|---:|\n")
        for k, v in agg.items():
            f.write(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot * 100:.1f} % |\n")


def full(src, dst):
    """src: a .ncu-rep, or the CSV `ncu -i x.ncu-rep --page raw --csv` made of it on the GPU box (reports are too big to bring back)"""
    if src.endswith(".csv"):
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n\n")
        f.write("| metric | unit | " + " | ".join(f"`{short_name(r[ki])}`" for r in rows[2:]) + " |\n")
        f.write("|---|---|" + "---:|" * (len(rows) - 2) + "\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |\n")


def short_name(full_name):
    """k_trace_cwbvh<(int)1, (bool)0>(rtb::TraceArgs) -> k_trace_cwbvh<1,0>; k_trace_cwbvh_frustum<...> -> k_trace_cwbvh_frustum"""
    import re
    base = full_name.split("(rtb::")[0].split("(const rtb::")[0]
    base = base[5:] if base.startswith("void ") else base
    m = re.match(r"(?:rtb::)?(\w+)<(.*)>", base)
    if not m:
        return base.replace("rtb::", "")
    name, args = m.group(1), re.sub(r"\((?:int|bool)\)", "", m.group(2)).replace(" ", "")
    return name if name.endswith(("frustum", "packet")) else f"{name}<{args}>"


def traffic(src_csv, workload, dst_json="profiles/trace_traffic.json"):
    """src_csv: `ncu -i x.ncu-rep --page raw --csv` of a --set full capture of scripts/profile_frame.py.  Adds, for `workload`,
    dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches of each kernel) to dst_json."""
    import json
    rows = list(csv.reader(open(src_csv)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        v = float(r[ri].replace(",", "")) * scale.get(units[ri], 1.0) + float(r[wi].replace(",", "")) * scale.get(units[wi], 1.0)
        agg.setdefault(short_name(r[ki]), []).append(v)
    data = json.load(open(dst_json))
    data[workload] = {k: int(sum(v) / len(v)) for k, v in agg.items()}
    data[workload]["_launches_captured"] = {k: len(v) for k, v in agg.items()}
    data[workload]["_capture"] = src_csv
    json.dump(data, open(dst_json, "w"), indent=1)
    print(workload, data[workload])


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
