#!/usr/bin/env python3
"""Shape of the deepest bundles of the bench workload: reads, graphs, nodes, transfrags, records."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
base, a = bench.build_workload(bench.PARTS, bench.PART_BUNDLES, 0)
nr = np.diff(a["b_read_off"].astype(np.int64))
order = np.argsort(-nr)
print("bundles", nr.size, "reads", nr.sum(), "top reads per bundle", nr[order[:12]].tolist())
print("bundles >= 100k reads:", int((nr >= 100000).sum()), ">= 16384:", int((nr >= 16384).sum()), "share of reads in top 10:", float(nr[order[:10]].sum() / nr.sum()))
span = (a["b_end"].astype(np.int64) - a["b_start"])[order[:5]]
print("span of top 5", span.tolist())
