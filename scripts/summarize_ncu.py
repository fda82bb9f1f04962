"""Summarise .ncu-rep captures (read here on the CPU box) into small text files under profiles/.

    python scripts/summarize_ncu.py gpurun_out/prof_render.ncu-rep profiles/r01_render_ncu.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, out):
    lines = []
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    ix = {h: i for i, h in enumerate(hdr)}
    lines.append(f"# ncu --set full --clock-control none --import-source on  ({rep})")
    lines.append(f"kernel: {vals[ix['Kernel Name']]}")
    for k in KEYS:
        if k in ix:
            lines.append(f"{k:75s} {vals[ix[k]]:>18s} {units[ix[k]]}")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    h2 = src[1]
    jx = {h: i for i, h in enumerate(h2)}
    data = [r for r in src[2:] if len(r) == len(h2)]
    stall = Counter()
    ops = Counter()
    ns = ninst = 0
    for r in data:
        try:
            n = int(r[jx["Instructions Executed"]])
            smp = int(r[jx["# Samples"]])
        except ValueError:
            continue
        ninst += n
        ns += smp
        toks = r[1].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        ops[op.split(".")[0]] += n
        for h in h2:
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    stall[h] += int(r[jx[h]])
                except ValueError:
                    pass
    lines.append(f"\nwarp-level instructions executed: {ninst}   stall samples: {ns}")
    lines.append("stall reasons (share of samples): " + ", ".join(f"{k[6:]} {100 * v / max(ns, 1):.1f}%" for k, v in stall.most_common(8)))
    lines.append("opcode mix (share of executed warp instructions): " + ", ".join(f"{k} {100 * v / max(ninst, 1):.1f}%" for k, v in ops.most_common(14)))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
