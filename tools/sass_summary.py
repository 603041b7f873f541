#!/usr/bin/env python3
"""Per-kernel instruction mix of libomni_b200.so from `cuobjdump -sass` -> profiles/sass_summary.txt (static counts: what the
code CONTAINS, not what runs).  Shows at a glance which kernels use the TMA engine (UTMALDG = bulk-tensor load, UBLKCP = bulk copy),
mbarriers (SYNCS), 256-bit stores, fp16x2 lanes, shuffles, shared-memory atomics, and that nothing uses tensor cores (no contraction
on this path).      python tools/sass_summary.py [LIB] [OUT]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "omnirevolve-image-processor_b200", "lib", "libomni_b200.so")
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "sass_summary.txt")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
GROUPS = [("UTMALDG", r"^UTMALDG"), ("UTMASTG", r"^UTMASTG"), ("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("UTCxMMA/LDTM", r"^(UTC|LDTM|STTM)"),
          ("HMMA/IMMA", r"^(HMMA|IMMA|QMMA|OMMA)"), ("STG.256", r"^STG\.E\.ENL2\.256"), ("STG", r"^STG"), ("LDG", r"^LDG"), ("LDS", r"^LDS"),
          ("STS", r"^STS"), ("ATOMS", r"^ATOMS"), ("ATOMG/RED", r"^(ATOMG|RED)"), ("SHFL", r"^SHFL"), ("VOTE/MATCH", r"^(VOTE|MATCH|VOTEU)"),
          ("LOP3", r"^(LOP3|ULOP3)"), ("SHF", r"^(SHF|USHF)"), ("PRMT", r"^PRMT"), ("IMAD/IADD", r"^(IMAD|IADD3|VIADD|UIADD3|UIMAD|LEA)"),
          ("HFMA2/HADD2", r"^(HFMA2|HADD2|HMUL2|HSET2|HSETP2|HMNMX2)"), ("FFMA/FADD/FMUL", r"^(FFMA|FADD|FMUL)"), ("POPC/FLO/BREV", r"^(POPC|FLO|BREV)"),
          ("BAR", r"^BAR")]
kern, rows = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        rows[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        rows[kern]["total"] += 1
        for name, rx in GROUPS:
            if re.match(rx, op):
                rows[kern][name] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(rows), capture_output=True, text=True).stdout.splitlines()
names = [g[0] for g in GROUPS]
with open(out, "w") as fh:
    fh.write("# static SASS instruction mix per kernel of %s (tools/sass_summary.py; cuobjdump -sass, sm_100a)\n" % os.path.basename(lib))
    fh.write("# kernels that contain TMA-engine instructions (UTMALDG / UTMASTG / UBLKCP): %s\n" % ", ".join(
        sorted({d.split("(")[0].replace("void ", "").split("<")[0] for d, c in zip(demangle, rows.values()) if c["UTMALDG"] + c["UTMASTG"] + c["UBLKCP"]})))
    fh.write("# tensor-core instructions anywhere (UTCxMMA / LDTM / STTM / HMMA / IMMA): %d (nothing on this path is a contraction)\n" %
             sum(c["UTCxMMA/LDTM"] + c["HMMA/IMMA"] for c in rows.values()))
    fh.write("%-64s %6s " % ("kernel", "total") + " ".join("%9s" % n[:9] for n in names) + "\n")
    for d, c in zip(demangle, rows.values()):
        short = d.split("(")[0].replace("void ", "")[:64]
        fh.write("%-64s %6d " % (short, c["total"]) + " ".join("%9d" % c[n] for n in names) + "\n")
print("wrote", out, len(rows), "kernels")
