"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep captured with
`--set full --import-source on` (kernels compiled with -lineinfo).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [min_share_pct] [launch_index]

Prints, per CUDA source line: share of warp instructions executed, share of stall samples, the line."""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          "regex:" + rx, "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    lines = []
    fname = ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            print("#", r[1])
            continue
        if r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or r[0] == "":
            continue
        try:
            ie = hdr["Instructions Executed"]
            lines.append((fname, int(r[0]), int(r[ie]), int(r[hdr["# Samples"]]), r[1]))
        except (ValueError, KeyError, IndexError):
            continue
    ti = sum(x[2] for x in lines) or 1
    ts = sum(x[3] for x in lines) or 1
    print(f"# total warp-instructions {ti}, stall samples {ts}")
    for f, ln, ins, smp, src in lines:
        if 100.0 * ins / ti >= min_pct or 100.0 * smp / ts >= min_pct:
            print(f"{f}:{ln:5d} inst {100.0 * ins / ti:5.1f}%  samples {100.0 * smp / ts:5.1f}%  | {src.strip()[:140]}")


if __name__ == "__main__":
    main()
