#!/usr/bin/env python3
"""SASS evidence for libstgpu.so (cuobjdump -sass):
 * every FFMA / DFMA belongs to the expansion of an IEEE division / square root (Newton steps behind a MUFU.RCP* / MUFU.RSQ*,
   or the out-of-line slow path $__internal_*): the library is built with -fmad=false, so no product-sum of the source may be
   contracted — a contracted multiply-add would change the bits of cov / abundance / flow against the reference's SSE build;
 * the kernels that move finished tiles use the bulk-copy path (UBLKCP = cp.async.bulk shared -> global).
usage: sass_check.py [libstgpu.so]   (exit 1 if a fused multiply-add has no division next to it)"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "stringtie_b200", "libstgpu.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    fn, label, window = None, "", []
    stats, loose, bulk = {}, [], {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn, label, window = m.group(1), "", []
            continue
        m = re.match(r"\s*(\$\S+|\.L_x_\d+):", line)
        if m:
            if m.group(1).startswith("$"):
                label = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if not m or fn is None:
            continue
        ins = m.group(1)
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        window.append(op)
        if len(window) > 96:
            window.pop(0)
        if op.startswith("UBLKCP"):
            bulk[fn] = bulk.get(fn, 0) + 1
        if op.startswith("FFMA") or op.startswith("DFMA"):
            near = any(w.startswith("MUFU.RCP") or w.startswith("MUFU.RSQ") for w in window)
            inside = label.startswith("$__internal") or ins.rstrip().endswith(", RZ")   # x * 2^64 + 0: the pre-scaling of a division, a plain product
            st = stats.setdefault(fn, [0, 0])
            st[0] += 1
            if not (near or inside):
                st[1] += 1
                loose.append((fn, ins))
    print("| kernel | FFMA / DFMA | not next to a division |")
    print("|---|---|---|")
    for k, (n, bad) in sorted(stats.items(), key=lambda x: -x[1][0]):
        print(f"| `{k[:90]}` | {n} | {bad} |")
    print()
    print("bulk copies (UBLKCP):", ", ".join(f"`{k[:60]}` x{n}" for k, n in bulk.items()) or "none")
    if loose:
        print("\nfused multiply-adds without a division in the 96 instructions before them:")
        for fn, ins in loose[:40]:
            print("  ", fn[:70], "|", ins)
    return 1 if loose else 0


if __name__ == "__main__":
    sys.exit(main())
