"""Hottest SASS instructions of one kernel in an ncu report (--set full --import-source on): stall samples, executions,
shared-memory wavefronts (actual / ideal).   python tools/ncu_sass_top.py REPORT.ncu-rep KERNEL_REGEX [N]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", "regex:" + rx,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break                                   # first function only
        hdr = r
        continue
    if hdr is not None and len(r) == len(hdr):
        data.append(r)
ix = {k: j for j, k in enumerate(hdr)}


def I(r, k):
    try:
        return int(r[ix[k]])
    except (ValueError, KeyError):
        return 0


tot = sum(I(r, "# Samples") for r in data)
tin = sum(I(r, "Instructions Executed") for r in data)
W, WI = sum(I(r, "L1 Wavefronts Shared") for r in data), sum(I(r, "L1 Wavefronts Shared Ideal") for r in data)
print(f"{len(data)} SASS lines, {tin} warp-instructions, {tot} stall samples, shared wavefronts {W} (ideal {WI})")
print("--- by shared-memory wavefronts: addr samples executed wavefronts ideal")
for r in sorted(data, key=lambda r: -I(r, "L1 Wavefronts Shared"))[:n]:
    if I(r, "L1 Wavefronts Shared"):
        print(r[ix["Address"]][-5:], I(r, "# Samples"), I(r, "Instructions Executed"), I(r, "L1 Wavefronts Shared"), I(r, "L1 Wavefronts Shared Ideal"), r[ix["Source"]][:90])
print("--- by stall samples: addr samples executed")
for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:n]:
    print(r[ix["Address"]][-5:], I(r, "# Samples"), I(r, "Instructions Executed"), r[ix["Source"]][:100])
