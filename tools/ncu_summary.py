#!/usr/bin/env python3
"""Markdown summary + per-launch averages of an `ncu --set full` capture of the closest-hit kernels.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv > profiles/rNN_extend_ncu_summary.md

The last lines (JSON) are the per-launch averages bench.py carries in NCU_CAPTURE: DRAM / L2 / L1 bytes per closest-hit launch,
averaged over the captured launches (the depth launches of one batch) exactly like bench.py averages the live launch duration.
"""
import csv
import json
import sys

ROWS = [
    ("gpu__time_duration.sum", "ms"),
    ("launch__registers_per_thread", ""),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "%"),
    ("smsp__inst_executed.sum", "inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "%"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "%"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "%"),
    ("l1tex__t_sector_hit_rate.pct", "%"),
    ("l1tex__t_bytes.sum", "GB"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sectors"),
    ("lts__t_sector_hit_rate.pct", "%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "%"),
    ("lts__t_bytes.sum", "GB"),
    ("dram__bytes_read.sum", "GB"),
    ("dram__bytes_write.sum", "GB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "lg_throttle", "branch_resolving", "not_selected",
          "dispatch_stall", "no_instruction", "mio_throttle", "barrier"]
SCALE = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "msecond": 1.0, "ms": 1.0, "usecond": 1e-3, "us": 1e-3, "second": 1e3, "s": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(name, r):
        if name not in col:
            return None
        v = data[r][col[name]].replace(",", "")
        try:
            f = float(v)
        except ValueError:
            return None
        u = units[col[name]]
        return f * SCALE.get(u, 1.0) if u in SCALE else f

    n = len(data)
    names = [data[r][col["Kernel Name"]] for r in range(n)]
    print("| metric | unit | " + " | ".join("d%d" % r for r in range(n)) + " |")
    print("|---|---|" + "---|" * n)
    print("| kernel | | " + " | ".join("`%s`" % nm.split("(")[0].replace("void ", "").replace("ptb::", "") for nm in names) + " |")
    for name, unit in ROWS:
        vals = [get(name, r) for r in range(n)]
        if all(v is None for v in vals):
            continue
        print("| `%s` | %s | " % (name, unit) + " | ".join("" if v is None else ("%.3g" % v) for v in vals) + " |")
    for s in STALLS:
        key = "smsp__average_warp_latency_issue_stalled_%s.ratio" % s
        alt = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
        k = key if key in col else alt
        vals = [get(k, r) for r in range(n)]
        if all(v is None for v in vals):
            continue
        print("| stall: %s (cycles per issued instruction) | | " % s + " | ".join("" if v is None else ("%.3g" % v) for v in vals) + " |")

    def total(name):
        vals = [get(name, r) for r in range(n)]
        return None if any(v is None for v in vals) else sum(vals)
    dur = total("gpu__time_duration.sum")
    wb_cycles = [get("l1tex__lsu_writeback_active_mem_lgds.sum", r) for r in range(n)]
    out = {
        "launches": n, "sum_duration_ms": dur,
        "dram_bytes_per_launch": (total("dram__bytes_read.sum") + total("dram__bytes_write.sum")) * 1e9 / n,
        "l2_bytes_per_launch": total("lts__t_bytes.sum") * 1e9 / n,
        "l1_tag_bytes_per_launch": total("l1tex__t_bytes.sum") * 1e9 / n if total("l1tex__t_bytes.sum") is not None else None,
        # L1 -> register-file write-back: cycles the LSU write-back port was busy x its 128 bytes per cycle (what a load INSTRUCTION costs
        # the data pipe whatever the lanes' addresses are)
        "l1_writeback_bytes_per_launch": None if any(v is None for v in wb_cycles) else sum(wb_cycles) * 128.0 / n,
        "time_weighted": {},
    }
    for name in ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                 "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
                 "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"):
        vals = [get(name, r) for r in range(n)]
        ds = [get("gpu__time_duration.sum", r) for r in range(n)]
        if all(v is not None for v in vals):
            out["time_weighted"][name] = sum(v * d for v, d in zip(vals, ds)) / sum(ds)
    print()
    print("```json")
    print(json.dumps(out, indent=1))
    print("```")


if __name__ == "__main__":
    main()
