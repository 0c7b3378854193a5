#!/usr/bin/env python
"""Per-kernel summary of an ncu report (one row per captured launch): the numbers profiles/ wants next to a bench line.
usage: tools/ncu_summary.py gpurun_out/X.ncu-rep > profiles/X_summary.txt"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time_us", 1e-3), ("dram__bytes_read.sum", "dram_rd_MB", None), ("dram__bytes_write.sum", "dram_wr_MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1), ("lts__t_sector_hit_rate.pct", "L2hit_%", 1),
        ("l1tex__t_sector_hit_rate.pct", "L1hit_%", 1), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", 1),
        ("launch__registers_per_thread", "regs", 1), ("smsp__inst_executed.sum", "Minst", 1e-6),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%", 1), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_%", 1)]
STALLS = ["long_scoreboard", "short_scoreboard", "no_instruction", "wait", "barrier", "math_pipe_throttle", "lg_throttle", "branch_resolving", "not_selected",
          "mio_throttle", "dispatch_stall", "imc_miss", "membar", "sleeping"]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * mult


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(data)} launches (ncu --set full, --clock-control none; times are single cold launches under replay)")
    print("%-46s" % "kernel" + " ".join("%10s" % c[1] for c in COLS))
    for r in data:
        name = r[ix["Kernel Name"]]
        name = name.replace("hkd::", "").split("(")[0][:46]
        out = []
        for key, label, scale in COLS:
            if key not in ix or r[ix[key]] in ("", "n/a"):
                out.append("%10s" % "-"); continue
            v = r[ix[key]].replace(",", "")
            if scale is None:
                out.append("%10.1f" % (to_bytes(v, units[ix[key]]) / 1e6))
            else:
                f = float(v) * scale
                if label == "time_us" and units[ix[key]] in ("usecond", "us"):
                    f = float(v)
                elif label == "time_us" and units[ix[key]] in ("msecond", "ms"):
                    f = float(v) * 1e3
                out.append("%10.2f" % f)
        print("%-46s" % name + " ".join(out))
    print("\n# warp stall reasons, warps per issue (smsp__average_warps_issue_stalled_<reason>_per_issue_active)")
    print("%-46s" % "kernel" + " ".join("%8s" % s[:8] for s in STALLS))
    for r in data:
        name = r[ix["Kernel Name"]].replace("hkd::", "").split("(")[0][:46]
        vals = []
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            vals.append("%8.2f" % float(r[ix[key]]) if key in ix and r[ix[key]] not in ("", "n/a") else "%8s" % "-")
        print("%-46s" % name + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
