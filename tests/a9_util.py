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


def split_engine_groups(G, a):
    """Per-bundle dicts (keys as orc.groups / split_reference_dump) from the batch-wide arrays of Engine.fetch_groups."""
    nb = a["b_start"].size
    ro, so, po, jo = (a[k].astype(np.int64) for k in ("b_read_off", "r_seg_off", "r_pair_off", "b_junc_off"))
    go, co, ao, rgo = (G[k].astype(np.int64) for k in ("GROUP_OFF", "COLOR_OFF", "APP_OFF", "RG_OFF"))
    out = []
    for b in range(nb):
        r0, r1 = int(ro[b]), int(ro[b + 1])
        s0, s1, p0, p1 = int(so[r0]), int(so[r1]), int(po[r0]), int(po[r1])
        g0, g1, c0, c1 = int(go[b]), int(go[b + 1]), int(co[b]), int(co[b + 1])
        j0, j1 = int(jo[b] + ao[b]), int(jo[b + 1] + ao[b + 1])
        ng = g1 - g0
        d = {"startgroup": G["STARTGROUP"][3 * b:3 * b + 3], "num_fragments": G["NUM_FRAGMENTS"][b], "frag_len": G["FRAG_LEN"][b],
             "color": c1 - c0,
             "r_strand": G["R_STRAND"][r0:r1], "pair_idx": G["PAIR_IDX"][p0:p1], "seg_start": G["SEG_START"][s0:s1],
             "seg_end": G["SEG_END"][s0:s1], "rjunc": G["RJUNC"][s0 - r0:s1 - r1],
             "g_alive": G["G_ALIVE"][g0:g1], "g_start": G["G_START"][g0:g1], "g_end": G["G_END"][g0:g1],
             "g_color": G["G_COLOR"][g0:g1], "g_next": G["G_NEXT"][g0:g1], "g_cov": G["G_COV"][g0:g1],
             "g_multi": G["G_MULTI"][g0:g1], "merge": G["G_MERGE"][g0:g1], "eqcol": G["EQCOL"][c0:c1],
             "rg_off": (rgo[r0:r1 + 1] - rgo[r0]).astype(np.uint32), "rg": G["RG"][int(rgo[r0]):int(rgo[r1])],
             "g_color2": G["G_COLOR2"][g0:g1], "g_neg_prop": G["G_NEG_PROP"][g0:g1], "eqcol2": G["EQCOL2"][c0:c1],
             "eqneg": G["EQNEG"][c0:c1], "eqpos": G["EQPOS"][c0:c1], "g2b": G["G2B"][3 * g0:3 * g1]}
        for k in J_KEYS:
            d["j_" + k] = G["J_" + k.upper()][j0:j1]
        nbun = G["N_BUN"][3 * b:3 * b + 3].astype(np.int64)
        nbn = G["N_BN"][3 * b:3 * b + 3].astype(np.int64)
        d["bun_off"] = np.concatenate([[0], np.cumsum(nbun)]).astype(np.int32)
        d["bn_off"] = np.concatenate([[0], np.cumsum(nbn)]).astype(np.int32)
        for k in ("len", "start", "last", "cov", "multi"):
            src = G["BUN_" + k.upper()]
            d["bun_" + k] = np.concatenate([src[3 * g0 + s * ng:3 * g0 + s * ng + int(nbun[s])] for s in range(3)])
        for k in ("start", "end", "cov", "next"):
            src = G["BN_" + k.upper()]
            d["bn_" + k] = np.concatenate([src[3 * g0 + s * ng:3 * g0 + s * ng + int(nbn[s])] for s in range(3)])
        out.append(d)
    return out
