"""Aggregate the per-SASS-instruction counters of an ncu report by CUDA source line.
    python profiles/ncu_source_hotspots.py report.ncu-rep [kernel-regex] [top]"""
import collections
import csv
import io
import subprocess
import sys


def main(path, kernel=None, top=30):
    cmd = ["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if kernel:
        cmd += ["--kernel-name", f"regex:{kernel}"]
    raw = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    h = rows[hdr]
    i_line, i_src, i_inst, i_stall = 0, 1, h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    agg = collections.OrderedDict()
    cur = None
    for r in rows[hdr + 1:]:
        if len(r) <= i_inst:
            continue
        if r[i_line].strip():
            cur = (r[i_line], r[i_src].strip())
            agg.setdefault(cur, [0, 0])
        if cur is None:
            continue
        try:
            agg[cur][0] += int(r[i_inst] or 0)
            agg[cur][1] += int(r[i_stall] or 0)
        except ValueError:
            pass
    tot_i = sum(v[0] for v in agg.values()) or 1
    tot_s = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {tot_i}, stall samples {tot_s}")
    for (line, src), (ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{line:>5s} {100 * ins / tot_i:5.1f}% inst {100 * st / tot_s:5.1f}% stall | {src[:120]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 30)
