"""Per-source-line instruction counts and stall samples of one ncu report (needs --import-source on, -lineinfo).
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
inst, samp, src, stalls = collections.Counter(), collections.Counter(), {}, collections.defaultdict(collections.Counter)
h = None
key = None
fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if "Instructions Executed" in r:
        h = r
        iI, iN, iA = h.index("Instructions Executed"), h.index("# Samples"), h.index("Address")
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if h is None or len(r) < len(h):
        continue
    if r[iA] == "-" or r[iA] == "":     # a source line row
        key = (fname, r[0])
        src[key] = r[1].strip()
        continue
    try:
        n, s = int(r[iI]), int(r[iN])
    except ValueError:
        continue
    inst[key] += n
    samp[key] += s
    for i, c in stall_cols:
        try:
            v = int(r[i])
        except ValueError:
            v = 0
        if v:
            stalls[key][c[6:]] += v
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp instructions", ti, "samples", ts)
for key, s in samp.most_common(top):
    st = ", ".join(f"{k}:{v}" for k, v in stalls[key].most_common(3))
    print(f"{key[0]}:{key[1]:>5} {100 * s / ts:5.1f}%smp {100 * inst[key] / ti:5.1f}%inst  [{st}]  {src.get(key, '')[:100]}")
