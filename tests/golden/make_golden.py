#!/usr/bin/env python3
"""Regenerates the committed golden fixtures from the REFERENCE ITSELF (needs /root/reference; run here, not on the
GPU box):  tools/gen_sam.py (seeded) -> oracle/_ref/ref_tool (unmodified reference main + dump hook, -p 1)
-> tests/golden/<name>.npz holding the packed-batch inputs that crossed infer_transcripts() and the reference's
stage outputs (bpcov + junction support after count_good_junctions, the junction table at the end of the junction
pass inside build_graphs with good_junc verdicts [a7_*], predictions at return).

  make -C oracle ref && python tests/golden/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from stringtie_b200.stgb import read_stgb  # noqa: E402

CASES = {
    # name: (gen_sam args, stringtie args)
    "short_s11": (["--loci", "24", "--pairs", "2500", "--seed", "11", "--long-intron-frac", "0.01"], []),
    "short_s12_deep": (["--loci", "6", "--pairs", "6000", "--seed", "12", "--long-intron-frac", "0.0",
                        "--nh-frac", "0.2"], []),
    # splice-site jitter + NM tags: junctions supported only by mismatchy reads -> the higherr / demotion pass
    "short_s13_jitter": (["--loci", "16", "--pairs", "4000", "--seed", "13", "--long-intron-frac", "0.02",
                          "--nh-frac", "0.1", "--jitter-frac", "0.06"], []),
}


def main():
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_tool")
    for name, (gargs, sargs) in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            sam, dump, gtf = (os.path.join(td, x) for x in ("in.sam", "dump.stgb", "out.gtf"))
            subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_sam.py"), "--sam", sam] + gargs)
            env = dict(os.environ, STG_DUMP=dump)
            subprocess.check_call([tool, sam, "-p", "1", "-o", gtf] + sargs, env=env)
            a = read_stgb(dump)
            out = os.path.join(ROOT, "tests", "golden", name + ".npz")
            np.savez_compressed(out, **a)
            print(name, "bundles", a["b_start"].size, "reads", a["r_start"].size, "cov floats", a["s1_cov"].size,
                  "preds", a["p_start"].size, "->", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
