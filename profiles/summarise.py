#!/usr/bin/env python3
"""profiles/summarise.py -- turn ncu outputs brought back in gpurun_out/ into the small tracked
summaries under profiles/.

  python profiles/summarise.py full  gpurun_out/prof_X.ncu-rep profiles/rNN_X_ncu.json  [note]
  python profiles/summarise.py list  gpurun_out/launches.csv   profiles/rNN_X_launches.json [note]

`full` reads one `ncu --set full` capture (first kernel in the report) through
`ncu -i ... --page raw --csv` and keeps the counters the roofline discussion needs;
`list` reads a `--metrics gpu__time_duration.sum --csv` launch list and reports each kernel's
share of the summed device time (cold-cache, serialised launches: compare SHARES, not absolutes).
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__lsu_writeback_active_mem_lg.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
]
STALL = "smsp__average_warps_issue_stalled_"
STALL2 = "smsp__average_warp_latency_issue_stalled_"


def full(rep, out, note):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = {"source": rep, "note": note, "kernels": []}
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h in ("Kernel Name", "Block Size", "Grid Size"):
                d[h] = v
            elif h in KEEP or "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                d[h] = {"value": v, "unit": u}
        try:
            t = d["gpu__time_duration.sum"]
            rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            d["dram_bytes_per_launch"] = float(rd["value"]) * scale[rd["unit"]] + float(wr["value"]) * scale[wr["unit"]]
            d["duration_us"] = float(t["value"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[t["unit"]]
        except Exception:
            pass
        res["kernels"].append(d)
    json.dump(res, open(out, "w"), indent=1)
    for k in res["kernels"]:
        print(k.get("Kernel Name", "?")[:70], k.get("duration_us"), "us", k.get("dram_bytes_per_launch"), "B")


def launch_list(path, out, note):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    res = {"source": path, "note": note, "total_us": tot,
           "kernels": [{"kernel": k, "launches": a[0], "total_us": a[1], "mean_us": a[1] / a[0], "share": a[1] / tot}
                       for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    json.dump(res, open(out, "w"), indent=1)
    for k in res["kernels"]:
        print("%-40s n=%4d mean=%9.1f us share=%.3f" % (k["kernel"][:40], k["launches"], k["mean_us"], k["share"]))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    (full if mode == "full" else launch_list)(src, dst, note)
