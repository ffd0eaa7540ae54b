#!/usr/bin/env python3
"""Times the integer form of ordered float32 sums (exact_chain.cuh, through stg_selftest_chain) against the bare chain of
dependent adds on list shapes of the C2 workload: one hot splice-site position, one hot graph node, many medium lists."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stringtie_b200.engine import Engine


def main():
    rng = np.random.default_rng(5)
    w = np.array([1.0, 0.5, 1.0 / 3.0, 0.25, 0.2], np.float32)
    p = np.array([0.93, 0.03, 0.015, 0.015, 0.01])
    cases = {}
    cases["one position, 1.3 M deposits of -w"] = [-w[rng.choice(5, 1300000, p=p)]]
    cases["one node, 3 M contributions bp*w"] = [(rng.integers(1, 151, 3000000).astype(np.float32) * w[rng.choice(5, 3000000, p=p)]).astype(np.float32)]
    cases["20 000 lists of 65..2000 deposits"] = [w[rng.choice(5, int(n), p=p)] for n in rng.integers(65, 2000, 20000)]
    cases["one position, +w / -w mixed around 0"] = [(w[rng.choice(5, 500000, p=p)] * rng.choice(np.array([1, -1], np.float32), 500000)).astype(np.float32)]
    eng = Engine(0)
    for name, xs in cases.items():
        off = np.zeros(len(xs) + 1, np.uint32); np.cumsum([x.size for x in xs], out=off[1:])
        x = np.concatenate(xs); s0 = np.zeros(len(xs), np.float32)
        res = {}
        for plain in (1, 0, 2):
            best = 1e9
            for _ in range(3):
                out, ms = eng.selftest_chain(off, x, s0, plain=plain, want_ms=True)
                best = min(best, ms)
            res[plain] = (out, best)
        same = np.array_equal(res[1][0].view(np.uint32), res[0][0].view(np.uint32)) and np.array_equal(res[1][0].view(np.uint32), res[2][0].view(np.uint32))
        print(f"{name}: plain {res[1][1]:.3f} ms, integer form: one warp {res[0][1]:.3f} ms, one CTA {res[2][1]:.3f} ms ({x.size / res[2][1] / 1e6:.2f} G terms/s), bits equal: {same}")
    eng.close()


if __name__ == "__main__":
    main()
