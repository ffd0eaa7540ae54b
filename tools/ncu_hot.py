#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build).
usage: tools/ncu_hot.py report.ncu-rep kernel_name [top]"""
import csv, subprocess, sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                                   "--kernel-name", kern], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    lines = []
    for r in rows:
        if hdr is None:
            if "# Samples" in r:
                hdr = r
                si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if len(r) <= max(si, ii) or r[0] == "":
            continue  # SASS rows have an empty line number: the CUDA row above already aggregates them
        try:
            lines.append((int(r[si]), int(r[ii]), r[0], r[1][:130]))
        except ValueError:
            continue
    tot_s = sum(x[0] for x in lines); tot_i = sum(x[1] for x in lines)
    print(f"kernel {kern}: {tot_s} samples, {tot_i} warp-level instructions executed")
    for s, n, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100 * s / max(tot_s, 1):5.1f}% samples {100 * n / max(tot_i, 1):5.1f}% inst | line {ln:>4} | {src.strip()}")


if __name__ == "__main__":
    main()
