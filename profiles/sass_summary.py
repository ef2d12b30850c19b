"""SASS evidence: per kernel of libyolo_b200.so, how often the Blackwell-specific instructions occur.
    python profiles/sass_summary.py > profiles/r02_sass_summary.txt
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM), UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk
(1-D bulk copy, either direction), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, WARPSYNC = convergence guards."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch_yolo_b200", "libyolo_b200.so")
WANT = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "SHFL", "VOTE", "WARPSYNC", "MUFU", "LDG", "STG", "ATOM", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur]["total"] += 1
            op = m.group(1)
            for w in WANT:
                if op.startswith(w):
                    kernels[cur][w] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (cuobjdump -sass, sm_100a)")
    print("# " + " ".join(f"{w:>8s}" for w in ["total"] + WANT) + "  kernel")
    for k, c in kernels.items():
        print("  " + " ".join(f"{c[w]:8d}" for w in ["total"] + WANT) + "  " + k)


if __name__ == "__main__":
    sys.exit(main())
