"""Helpers for SURVEY §8 rows a8 + a9 (first half): per-bundle views of the reference's dump taken after the read loop of
build_graphs (arrays a9_* written by oracle/ref_hook.cpp) and the comparison with another implementation's result."""
import numpy as np

J_KEYS = ("start", "end", "strand", "nreads", "nreads_good", "left", "right", "nm", "mm")
GROUP_KEYS = ("g_start", "g_end", "g_color", "g_next", "g_cov", "g_multi", "g_color2", "g_neg_prop")


def split_reference_dump(ref, a):
    nb = a["b_start"].size
    ro, so, po = (a[k].astype(np.int64) for k in ("b_read_off", "r_seg_off", "r_pair_off"))
    out = []
    for b in range(nb):
        g0, g1 = int(ref["a9_group_off"][b]), int(ref["a9_group_off"][b + 1])
        c0, c1 = int(ref["a9_color_off"][b]), int(ref["a9_color_off"][b + 1])
        j0, j1 = int(ref["a9_junc_off"][b]), int(ref["a9_junc_off"][b + 1])
        r0, r1 = int(ro[b]), int(ro[b + 1])
        s0, s1, p0, p1 = int(so[r0]), int(so[r1]), int(po[r0]), int(po[r1])
        rg0, rg1 = int(ref["a9_rg_off"][r0]), int(ref["a9_rg_off"][r1])
        d = {"startgroup": ref["a9_startgroup"][3 * b:3 * b + 3], "num_fragments": ref["a9_num_fragments"][b],
             "frag_len": ref["a9_frag_len"][b], "color": ref["a9_color"][b],
             "r_strand": ref["a9_r_strand"][r0:r1], "pair_idx": ref["a9_pair_idx"][p0:p1],
             "seg_start": ref["a9_seg_start"][s0:s1], "seg_end": ref["a9_seg_end"][s0:s1],
             "rjunc": ref["a9_rjunc"][s0 - r0:s1 - r1],
             "g_alive": ref["a9_g_alive"][g0:g1], "g_start": ref["a9_g_start"][g0:g1], "g_end": ref["a9_g_end"][g0:g1],
             "g_color": ref["a9_g_color"][g0:g1], "g_next": ref["a9_g_next"][g0:g1], "g_cov": ref["a9_g_cov"][g0:g1],
             "g_multi": ref["a9_g_multi"][g0:g1], "merge": ref["a9_merge"][g0:g1], "eqcol": ref["a9_eqcol"][c0:c1],
             "rg_off": ref["a9_rg_off"][r0:r1 + 1] - ref["a9_rg_off"][r0], "rg": ref["a9_rg"][rg0:rg1]}
        for k in J_KEYS:
            d["j_" + k] = ref["a9_j_" + k][j0:j1]
        if "a9b_g_color" in ref and ref["a9b_bun_off"].size == 3 * nb + 1:   # second half dumped as well
            bo, no = ref["a9b_bun_off"].astype(np.int64), ref["a9b_bn_off"].astype(np.int64)
            d.update({"g_color2": ref["a9b_g_color"][g0:g1], "g_neg_prop": ref["a9b_g_neg_prop"][g0:g1],
                      "eqcol2": ref["a9b_eqcol"][c0:c1], "eqneg": ref["a9b_eqneg"][c0:c1], "eqpos": ref["a9b_eqpos"][c0:c1],
                      "bun_off": (bo[3 * b:3 * b + 4] - bo[3 * b]).astype(np.int32),
                      "bn_off": (no[3 * b:3 * b + 4] - no[3 * b]).astype(np.int32),
                      "g2b": ref["a9b_g2b"][3 * g0:3 * g1]})
            for k in ("len", "start", "last", "cov", "multi"):
                d["bun_" + k] = ref["a9b_bun_" + k][bo[3 * b]:bo[3 * b + 3]]
            for k in ("start", "end", "cov", "next"):
                d["bn_" + k] = ref["a9b_bn_" + k][no[3 * b]:no[3 * b + 3]]
        out.append(d)
    return out


def assert_same(got, want):
    """Exact equality, float32 compared by bit pattern; fields of groups merged away (dead slots) are not compared."""
    assert len(got) == len(want)
    for b, (g, w) in enumerate(zip(got, want)):
        alive = np.asarray(w["g_alive"]).astype(bool)
        for k in w:
            gv, wv = np.asarray(g[k]), np.asarray(w[k])
            assert gv.shape == wv.shape, f"bundle {b}, {k}: shape {gv.shape} vs {wv.shape}"
            if k in GROUP_KEYS:
                gv, wv = gv[alive], wv[alive]
            if gv.dtype == np.float32:
                gv, wv = gv.view(np.uint32), wv.view(np.uint32)
            bad = np.nonzero(gv != wv)[0] if gv.ndim else (np.array([0]) if gv != wv else np.array([], dtype=int))
            assert bad.size == 0, f"bundle {b}, {k}: {bad.size} entries differ, first at {bad[:5]}"
