"""Per-source-line instruction and stall-sample shares of one kernel from a .ncu-rep captured with
--import-source on (read on the CPU box):

    python scripts/ncu_lines.py gpurun_out/prof_render.ncu-rep profiles/r01_render_lines.txt [min_share_percent]
"""
import csv
import io
import subprocess
import sys


def main(rep, out, min_share=0.15):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, {}
    for r in csv.reader(io.StringIO(txt)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
            jn, js = hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif hdr and len(r) == len(hdr) and r[0] != "":
            try:
                n, smp = int(r[jn]), int(r[js])
            except ValueError:
                continue
            a = agg.setdefault((cur, int(r[0])), [0, 0, r[1].strip()[:110]])
            a[0] += n
            a[1] += smp
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    lines = [f"# {rep}: warp-level instructions executed {tot}, stall samples {ts}",
             "# per file:  share of instructions / share of stall samples"]
    byfile = {}
    for (f, _), (n, s, _) in agg.items():
        b = byfile.setdefault(f, [0, 0])
        b[0] += n
        b[1] += s
    for f, (n, s) in sorted(byfile.items(), key=lambda x: -x[1][0]):
        lines.append(f"{f:28s} {100 * n / tot:6.2f}%  {100 * s / ts:6.2f}%")
    lines.append(f"# per line (>= {min_share} % of instructions): file:line  Minstr  inst%  samples%  source")
    for (f, l), (n, s, src) in sorted(agg.items(), key=lambda x: (x[0][0], x[0][1])):
        if 100 * n / tot >= min_share:
            lines.append(f"{f}:{l:<4d} {n / 1e6:9.1f} {100 * n / tot:6.2f} {100 * s / ts:6.2f}  {src}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.15)
