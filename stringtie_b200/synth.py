"""Seeded synthetic packed batches (SURVEY.md §8d, config C2 shape) generated directly in the boundary schema.

What the generator emits is what the reference's producer (processRead, rlink.cpp:659-983) would have left in
BundleData for coordinate-sorted 2x150 paired-end spliced alignments: reads in start order per bundle, a junction
table sorted unique by (start, end, strand desc) behind the null sentinel row (rlink.cpp:821-824), per-junction
nreads / nm / mm sums (rlink.cpp:913-929), symmetric mate links (rlink.cpp:942-971), XS-derived strand on spliced
reads only, and NH-derived weights 1/NH (rlink.cpp:899).  Everything is vectorised numpy so that 10^7-10^8
alignments are practical; the per-bundle mean is ~1 700 alignments over a ~30 kb span as in C2.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

MAXSEG = 6


def make_batch(n_bundles: int = 64, reads_per_bundle: float = 1700.0, seed: int = 20251019, read_len: int = 150,
               nh_probs=(0.93, 0.03, 0.015, 0.015, 0.01), chrom_gap: int = 1000, expr_sigma: float = 2.0,
               unpaired_frac: float = 0.03, max_intron: int = 60000, mean_exons: float = 9.0) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    RL = read_len
    nb = n_bundles
    # ---- gene models: one locus per bundle, up to 4 isoforms by exon skipping --------------------------------
    nex = np.clip(1 + rng.geometric(1.0 / mean_exons, nb), 1, 24)
    ex_off = np.zeros(nb + 1, dtype=np.int64)
    np.cumsum(nex, out=ex_off[1:])
    NE = int(ex_off[-1])
    ex_bundle = np.repeat(np.arange(nb), nex)
    ex_len = np.clip(np.exp(rng.normal(np.log(150.0), 0.8, NE)), 30, 5000).astype(np.int64)
    in_len = np.clip(np.exp(rng.normal(np.log(1500.0), 1.3, NE)), 70, max_intron).astype(np.int64)
    first = np.zeros(NE, dtype=bool)
    first[ex_off[:-1]] = True
    in_len[first] = 0  # the "intron" before the first exon
    # within-bundle exon starts (relative to the locus start)
    step = ex_len + in_len
    cs = np.cumsum(step)
    base = np.repeat(cs[ex_off[:-1]] - step[ex_off[:-1]], nex)
    ex_rel_start = cs - ex_len - base + 0  # start of exon relative to locus start (first exon at 0)
    # isoforms: mask[iso, exon]; isoform 0 keeps everything, others skip internal exons with p=0.3
    NISO = 4
    keep = rng.random((NISO, NE)) > 0.3
    keep[0, :] = True
    last = np.zeros(NE, dtype=bool)
    last[ex_off[1:] - 1] = True
    keep[:, first] = True
    keep[:, last] = True
    # transcript-coordinate cumulative lengths per isoform
    iso_cum = np.zeros((NISO, NE), dtype=np.int64)  # exclusive prefix of kept exon lengths within the bundle
    iso_len = np.zeros((NISO, nb), dtype=np.int64)
    for i in range(NISO):
        l = np.where(keep[i], ex_len, 0)
        c = np.cumsum(l)
        bstart = np.repeat(c[ex_off[:-1]] - l[ex_off[:-1]], nex)
        iso_cum[i] = c - l - bstart
        iso_len[i] = np.add.reduceat(l, ex_off[:-1])
    strand_gene = rng.choice(np.array([-1, 1], dtype=np.int8), nb)
    # ---- fragments -----------------------------------------------------------------------------------------
    expr = np.exp(rng.normal(0.0, expr_sigma, nb))
    nfrag_total = int(nb * reads_per_bundle / 2)
    nfrag = np.maximum(1, (expr / expr.sum() * nfrag_total).astype(np.int64))
    F = int(nfrag.sum())
    f_bundle = np.repeat(np.arange(nb), nfrag)
    f_iso = rng.integers(0, NISO, F)
    tl = iso_len[f_iso, f_bundle]
    short = tl < RL + 10
    f_iso = np.where(short, 0, f_iso)
    tl = iso_len[f_iso, f_bundle]
    ok = tl >= RL + 10
    f_bundle, f_iso, tl = f_bundle[ok], f_iso[ok], tl[ok]
    F = f_bundle.size
    flen = np.minimum(tl, np.maximum(RL, rng.normal(300, 50, F).astype(np.int64)))
    tstart = (rng.random(F) * (tl - flen + 1)).astype(np.int64)
    nh = rng.choice(np.arange(1, len(nh_probs) + 1), F, p=np.array(nh_probs) / np.sum(nh_probs)).astype(np.int16)
    unpaired = rng.random(F) < unpaired_frac

    def map_read(t0):
        """transcript interval [t0, t0+RL) of each fragment -> up to MAXSEG genomic segments (relative coords)."""
        seg_s = np.zeros((MAXSEG, F), dtype=np.int64)
        seg_e = np.zeros((MAXSEG, F), dtype=np.int64)
        nseg = np.zeros(F, dtype=np.int64)
        # exon containing t0: search the isoform's cumulative array inside the bundle
        lo = ex_off[f_bundle]
        hi = ex_off[f_bundle + 1]
        pos = t0.copy()
        remaining = np.full(F, RL, dtype=np.int64)
        # linear walk over exons (at most 24 per bundle), vectorised over fragments
        e = lo.copy()
        # advance e to the exon containing pos
        for _ in range(24):
            cum = iso_cum[f_iso, np.minimum(e, NE - 1)]
            ln = np.where(keep[f_iso, np.minimum(e, NE - 1)], ex_len[np.minimum(e, NE - 1)], 0)
            adv = (e < hi - 1) & (pos >= cum + ln)
            e = np.where(adv, e + 1, e)
        for _ in range(24 + MAXSEG):
            ee = np.minimum(e, NE - 1)
            cum = iso_cum[f_iso, ee]
            ln = np.where(keep[f_iso, ee], ex_len[ee], 0)
            inside = (remaining > 0) & (e < hi) & (ln > 0) & (pos >= cum) & (pos < cum + ln) & (nseg < MAXSEG)
            take = np.minimum(remaining, cum + ln - pos)
            gs = ex_rel_start[ee] + (pos - cum)
            idx = np.where(inside)[0]
            seg_s[nseg[idx], idx] = gs[idx]
            seg_e[nseg[idx], idx] = gs[idx] + take[idx] - 1
            nseg[idx] += 1
            remaining = np.where(inside, remaining - take, remaining)
            pos = np.where(inside, pos + take, pos)
            e = np.where((remaining > 0) & (e < hi), e + 1, e)
        return seg_s, seg_e, nseg

    s1, e1, n1 = map_read(tstart)
    s2, e2, n2 = map_read(tstart + flen - RL)
    # bundle genomic placement: consecutive loci separated by chrom_gap (> runoffdist=200, stringtie.cpp:146)
    span = np.add.reduceat(step, ex_off[:-1])  # exonic+intronic length of each locus
    b_start = np.zeros(nb, dtype=np.int64)
    b_start[0] = 10000
    b_start[1:] = 10000 + np.cumsum(span[:-1] + chrom_gap)
    # ---- reads: two per fragment (or one when unpaired) ----------------------------------------------------
    keep2 = ~unpaired
    R1 = F
    R2 = int(keep2.sum())
    NR = R1 + R2
    r_frag = np.concatenate([np.arange(F), np.arange(F)[keep2]])
    r_nseg = np.concatenate([n1, n2[keep2]])
    segs_s = np.concatenate([s1, s2[:, keep2]], axis=1) + b_start[f_bundle][r_frag][None, :]
    segs_e = np.concatenate([e1, e2[:, keep2]], axis=1) + b_start[f_bundle][r_frag][None, :]
    r_bundle = f_bundle[r_frag]
    r_startpos = segs_s[0]
    # sort by (bundle, start): the order of a coordinate-sorted BAM
    order = np.lexsort((r_startpos, r_bundle))
    inv = np.empty(NR, dtype=np.int64)
    inv[order] = np.arange(NR)
    r_frag, r_nseg, r_bundle = r_frag[order], r_nseg[order], r_bundle[order]
    segs_s, segs_e = segs_s[:, order], segs_e[:, order]
    r_nh = nh[r_frag]
    w = (np.float32(1.0) / r_nh.astype(np.float32)).astype(np.float32)
    spliced = r_nseg > 1
    r_strand = np.where(spliced, strand_gene[r_bundle], 0).astype(np.int8)
    b_read_off = np.zeros(nb + 1, dtype=np.int64)
    np.cumsum(np.bincount(r_bundle, minlength=nb), out=b_read_off[1:])
    # flatten segments
    seg_off = np.zeros(NR + 1, dtype=np.int64)
    np.cumsum(r_nseg, out=seg_off[1:])
    NS = int(seg_off[-1])
    seg_read = np.repeat(np.arange(NR), r_nseg)
    seg_k = np.arange(NS) - seg_off[seg_read]
    seg_start = segs_s[seg_k, seg_read]
    seg_end = segs_e[seg_k, seg_read]
    r_end = seg_end[seg_off[1:] - 1]
    r_len = np.add.reduceat(seg_end - seg_start + 1, seg_off[:-1])
    # mate links (symmetric), bundle-local indices
    mate_pos = np.full(NR, -1, dtype=np.int64)
    first_idx = inv[np.arange(F)[keep2]]
    second_idx = inv[F + np.arange(R2)]
    mate_pos[first_idx] = second_idx
    mate_pos[second_idx] = first_idx
    has_mate = mate_pos >= 0
    pair_off = np.zeros(NR + 1, dtype=np.int64)
    np.cumsum(has_mate.astype(np.int64), out=pair_off[1:])
    pair_idx = (mate_pos[has_mate] - b_read_off[r_bundle[has_mate]]).astype(np.int32)
    pair_cnt = w[has_mate]
    # ---- junction table -------------------------------------------------------------------------------------
    is_intron = seg_k > 0  # intron i sits before segment i
    jr = seg_read[is_intron]
    js = seg_end[np.where(is_intron)[0] - 1]
    je = seg_start[is_intron]
    jb = r_bundle[jr]
    jst = r_strand[jr]
    key_sort = np.lexsort((-jst.astype(np.int64), je, js, jb))
    jb_s, js_s, je_s, jst_s = jb[key_sort], js[key_sort], je[key_sort], jst[key_sort]
    newj = np.ones(jb_s.size, dtype=bool)
    if jb_s.size > 1:
        newj[1:] = (jb_s[1:] != jb_s[:-1]) | (js_s[1:] != js_s[:-1]) | (je_s[1:] != je_s[:-1]) | (jst_s[1:] != jst_s[:-1])
    uid_sorted = np.cumsum(newj) - 1  # unique junction id in (bundle, start, end, strand desc) order
    NJU = int(uid_sorted[-1]) + 1 if jb_s.size else 0
    u_bundle = jb_s[newj]
    # rows per bundle: sentinel + its unique junctions (only bundles that have junctions get a sentinel)
    ju_count = np.bincount(u_bundle, minlength=nb) if NJU else np.zeros(nb, dtype=np.int64)
    rows = np.where(ju_count > 0, ju_count + 1, 0)
    b_junc_off = np.zeros(nb + 1, dtype=np.int64)
    np.cumsum(rows, out=b_junc_off[1:])
    NJ = int(b_junc_off[-1])
    j_start = np.zeros(NJ, dtype=np.uint32)
    j_end = np.zeros(NJ, dtype=np.uint32)
    j_strand = np.zeros(NJ, dtype=np.int8)
    j_nreads = np.zeros(NJ, dtype=np.float64)
    j_nm = np.zeros(NJ, dtype=np.float64)
    j_mm = np.zeros(NJ, dtype=np.float64)
    rjunc_idx = np.zeros(NS - NR, dtype=np.int32)
    if NJU:
        u_first = np.zeros(nb + 1, dtype=np.int64)
        np.cumsum(ju_count, out=u_first[1:])
        u_local = np.arange(NJU) - u_first[u_bundle] + 1  # +1: sentinel row
        u_row = b_junc_off[u_bundle] + u_local
        j_start[u_row] = js_s[newj]
        j_end[u_row] = je_s[newj]
        j_strand[u_row] = jst_s[newj]
        # per-intron local row, back in read order
        local_sorted = u_local[uid_sorted]
        local = np.empty(jb.size, dtype=np.int64)
        local[key_sort] = local_sorted
        rjunc_idx[:] = local.astype(np.int32)
        wj = w[jr].astype(np.float64)
        row = b_junc_off[jb] + local
        np.add.at(j_nreads, row, wj)
        np.add.at(j_nm, row, np.where(r_nh[jr] > 2, wj, 0.0))
        intron_pos = np.where(is_intron)[0]
        lenl = seg_end[intron_pos - 1] - seg_start[intron_pos - 1] + 1
        lenr = seg_end[intron_pos] - seg_start[intron_pos] + 1
        np.add.at(j_mm, row, np.where((lenl > 25) & (lenr > 25), wj, 0.0))
    b_end = np.zeros(nb, dtype=np.int64)
    np.maximum.at(b_end, r_bundle, r_end)
    b_first = np.full(nb, np.iinfo(np.int64).max)
    np.minimum.at(b_first, r_bundle, segs_s[0])
    empty = b_read_off[1:] == b_read_off[:-1]
    b_first[empty] = b_start[empty]
    b_end[empty] = b_start[empty] + 10
    return {
        "b_start": b_first.astype(np.int32), "b_end": b_end.astype(np.int32),
        "b_read_off": b_read_off.astype(np.uint32), "b_junc_off": b_junc_off.astype(np.uint32),
        "r_start": segs_s[0].astype(np.uint32), "r_end": r_end.astype(np.uint32), "r_len": r_len.astype(np.uint32),
        "r_count": w, "r_nh": r_nh.astype(np.int16), "r_strand": r_strand,
        "r_flags": np.zeros(NR, dtype=np.uint8), "r_seg_off": seg_off.astype(np.uint32),
        "r_pair_off": pair_off.astype(np.uint32), "seg_start": seg_start.astype(np.uint32),
        "seg_end": seg_end.astype(np.uint32), "rjunc_idx": rjunc_idx, "pair_idx": pair_idx,
        "pair_cnt": pair_cnt.astype(np.float32), "j_start": j_start, "j_end": j_end, "j_strand": j_strand,
        "j_guide_match": np.zeros(NJ, dtype=np.int8), "j_nreads": j_nreads, "j_nm": j_nm, "j_mm": j_mm,
    }


def make_adversarial(seed: int = 1, n_bundles: int = 12, max_reads: int = 400,
                     neg_links: bool = False, spans=None) -> Dict[str, np.ndarray]:
    """Small random batches built to stress the kernels rather than to look like RNA-seq: spans that straddle
    the 1024-position sub-tiles and the BSIZE=10000 blocks, inexact weights (1/3, 1/7, 0.1 ...), many deposits on
    one position, overlapping / nested / opposite-strand / unstranded mates, unpaired remainders, negative
    pair_idx (only with neg_links=True: the reference itself indexes rd[-1] there, rlink.cpp:151-152, so such
    inputs can only be checked against the C oracle), unitig reads, reads with many segments, empty bundles.
    `spans`: bundle lengths to draw from (long spans with few reads give bundles of many separate groups)."""
    rng = np.random.default_rng(seed)
    out = {k: [] for k in ("b_start", "b_end", "r_start", "r_end", "r_len", "r_count", "r_nh", "r_strand", "r_flags",
                           "seg_start", "seg_end", "rjunc_idx", "pair_idx", "pair_cnt", "j_start", "j_end", "j_strand",
                           "j_guide_match", "j_nreads", "j_nm", "j_mm")}
    b_read_off, b_junc_off, r_seg_off, r_pair_off = [0], [0], [0], [0]
    weights = np.array([1.0, 0.5, 1.0 / 3.0, 0.25, 0.2, 1.0 / 7.0, 0.1, 2.0, 3.0, 1.0 / 3.0 + 1.0 / 3.0], dtype=np.float32)
    pos0 = 5000
    for b in range(n_bundles):
        span = int(rng.choice(spans if spans is not None else [40, 900, 1023, 1024, 1025, 2048, 3000, 9999, 10000, 10001, 20500, 31000]))
        nreads = 0 if b == 3 else int(rng.integers(1, max_reads))
        start = pos0
        end = pos0 + span - 1
        pos0 = end + 1000
        hot = rng.integers(start, end + 1, size=6)  # positions many reads share
        reads = []
        for n in range(nreads):
            nseg = int(rng.choice([1, 1, 1, 2, 2, 3, 4, 8, 30])) if span > 400 else int(rng.choice([1, 2]))
            # segment boundaries: sorted distinct points inside the bundle
            pts = np.sort(rng.choice(np.arange(start, end + 1), size=min(2 * nseg, span), replace=False))
            if pts.size < 2:
                pts = np.array([start, end])
            if pts.size % 2:
                pts = pts[:-1]
            if rng.random() < 0.5:
                pts[rng.integers(0, pts.size)] = hot[rng.integers(0, hot.size)]
                pts = np.sort(pts)
                # remove duplicates that would make start > end ordering ambiguous
                pts = np.unique(pts)
                if pts.size % 2:
                    pts = pts[:-1]
                if pts.size < 2:
                    pts = np.array([start, min(start + 5, end)])
            segs = [(int(pts[2 * i]), int(pts[2 * i + 1])) for i in range(pts.size // 2)]
            # segments must be separated by at least one base (an intron)
            clean = [segs[0]]
            for s, e in segs[1:]:
                if s > clean[-1][1] + 1:
                    clean.append((s, e))
            reads.append(clean)
        reads.sort(key=lambda sg: sg[0][0])
        # junction table
        strands = rng.choice(np.array([-1, 0, 1], dtype=np.int8), size=nreads, p=[0.4, 0.2, 0.4])
        jset = {}
        for n, sg in enumerate(reads):
            for i in range(1, len(sg)):
                jset.setdefault((sg[i - 1][1], sg[i][0], int(strands[n])), None)
        jkeys = sorted(jset.keys(), key=lambda k: (k[0], k[1], -k[2]))
        rows = {k: i + 1 for i, k in enumerate(jkeys)}
        nj = len(jkeys) + (1 if jkeys else 0)
        jn = np.zeros(nj)
        jnm = np.zeros(nj)
        jmm = np.zeros(nj)
        counts = weights[rng.integers(0, weights.size, size=nreads)] * rng.choice([1, 1, 1, 2, 5], size=nreads).astype(np.float32)
        nhs = rng.choice(np.array([1, 1, 2, 3, 5], dtype=np.int16), size=nreads)
        flags = (rng.random(nreads) < 0.03).astype(np.uint8)  # unitig
        # mates: pair up random reads (i<j), some reads get two links, some links negative
        links = [[] for _ in range(nreads)]
        for _ in range(nreads // 2):
            i, j = sorted(rng.integers(0, nreads, size=2).tolist()) if nreads > 1 else (0, 0)
            if i == j:
                continue
            c = np.float32(min(counts[i], counts[j]) * rng.choice([1.0, 0.5, 1.0 / 3.0]))
            links[i].append((j, c))
            links[j].append((i, c))
        if neg_links and nreads > 2:
            links[int(rng.integers(0, nreads))].append((-1, np.float32(0.25)))
        for n, sg in enumerate(reads):
            out["r_start"].append(sg[0][0]); out["r_end"].append(sg[-1][1])
            out["r_len"].append(sum(e - s + 1 for s, e in sg)); out["r_count"].append(counts[n])
            out["r_nh"].append(nhs[n]); out["r_strand"].append(strands[n]); out["r_flags"].append(flags[n])
            for s, e in sg:
                out["seg_start"].append(s); out["seg_end"].append(e)
            for i in range(1, len(sg)):
                r = rows[(sg[i - 1][1], sg[i][0], int(strands[n]))]
                out["rjunc_idx"].append(r)
                jn[r] += counts[n]
                if nhs[n] > 2:
                    jnm[r] += counts[n]
                if rng.random() < 0.7:
                    jmm[r] += counts[n]
            for j, c in links[n]:
                out["pair_idx"].append(j); out["pair_cnt"].append(c)
            r_seg_off.append(len(out["seg_start"])); r_pair_off.append(len(out["pair_idx"]))
        if jkeys:
            out["j_start"].append(0); out["j_end"].append(0); out["j_strand"].append(0)
            for k in jkeys:
                out["j_start"].append(k[0]); out["j_end"].append(k[1]); out["j_strand"].append(k[2])
            # make some junctions "weak" (nreads - nm < junctionthr, mm == 0) to hit the 25-bp anchor rule
            weak = rng.random(nj) < 0.3
            jnm = np.where(weak, jn, jnm)
            jmm = np.where(weak, 0.0, jmm)
            out["j_nreads"].extend(jn.tolist()); out["j_nm"].extend(jnm.tolist()); out["j_mm"].extend(jmm.tolist())
            out["j_guide_match"].extend([0] * nj)
        out["b_start"].append(start); out["b_end"].append(end)
        b_read_off.append(len(out["r_start"])); b_junc_off.append(len(out["j_start"]))
    from .stgb import INPUT_FIELDS
    res = {k: np.asarray(v, dtype=INPUT_FIELDS[k]) for k, v in out.items()}
    res["b_read_off"] = np.asarray(b_read_off, dtype=np.uint32)
    res["b_junc_off"] = np.asarray(b_junc_off, dtype=np.uint32)
    res["r_seg_off"] = np.asarray(r_seg_off, dtype=np.uint32)
    res["r_pair_off"] = np.asarray(r_pair_off, dtype=np.uint32)
    return res


def make_sparse(n_reads: int = 3000, seed: int = 1, n_bundles: int = 2) -> Dict[str, np.ndarray]:
    """Bundles of short unspliced reads spaced further apart than the grouping distance: every read (or mate pair)
    founds its own group, so a bundle holds thousands of groups.  Odd bundles are paired (mates 3 reads apart)."""
    from .stgb import INPUT_FIELDS
    rng = np.random.default_rng(seed)
    out = {k: [] for k in ("b_start", "b_end", "r_start", "r_end", "r_len", "r_count", "r_nh", "r_strand", "r_flags",
                           "seg_start", "seg_end", "rjunc_idx", "pair_idx", "pair_cnt", "j_start", "j_end", "j_strand",
                           "j_guide_match", "j_nreads", "j_nm", "j_mm")}
    b_read_off, b_junc_off, r_seg_off, r_pair_off = [0], [0], [0], [0]
    pos0 = 20000
    for b in range(n_bundles):
        starts = pos0 + np.cumsum(rng.integers(130, 400, size=n_reads))
        lens = rng.integers(30, 60, size=n_reads)
        counts = rng.choice(np.array([1.0, 0.5, 1.0 / 3.0], dtype=np.float32), size=n_reads)
        if b % 2 == 0:   # every third position carries a second, shorter read inside the first one's group
            keep = np.repeat(np.arange(n_reads), np.where(np.arange(n_reads) % 3 == 0, 2, 1))
            first = np.concatenate([[True], keep[1:] != keep[:-1]])
            starts, lens, counts = starts[keep] + np.where(first, 0, 5), np.where(first, lens[keep], lens[keep] - 10), counts[keep]
        nr = starts.size
        for n in range(nr):
            out["r_start"].append(int(starts[n])); out["r_end"].append(int(starts[n] + lens[n] - 1))
            out["r_len"].append(int(lens[n])); out["r_count"].append(counts[n]); out["r_nh"].append(int(rng.choice([1, 1, 2])))
            out["r_strand"].append(0); out["r_flags"].append(0)
            out["seg_start"].append(int(starts[n])); out["seg_end"].append(int(starts[n] + lens[n] - 1))
            if b % 2 == 1:   # mates: n <-> n + 3 within blocks of 6
                m = n + 3 if n % 6 < 3 else n - 3
                if 0 <= m < nr:
                    out["pair_idx"].append(m); out["pair_cnt"].append(np.float32(min(counts[n], counts[m])))
            r_seg_off.append(len(out["seg_start"])); r_pair_off.append(len(out["pair_idx"]))
        out["b_start"].append(pos0); out["b_end"].append(int(starts[-1] + 100))
        pos0 = int(starts[-1]) + 100000
        b_read_off.append(len(out["r_start"])); b_junc_off.append(0)
    res = {k: np.asarray(v, dtype=INPUT_FIELDS[k]) for k, v in out.items()}
    res["b_read_off"] = np.asarray(b_read_off, dtype=np.uint32)
    res["b_junc_off"] = np.asarray(b_junc_off, dtype=np.uint32)
    res["r_seg_off"] = np.asarray(r_seg_off, dtype=np.uint32)
    res["r_pair_off"] = np.asarray(r_pair_off, dtype=np.uint32)
    return res


def replicate(a: Dict[str, np.ndarray], k: int, gap: int = 100000) -> Dict[str, np.ndarray]:
    """k coordinate-shifted copies of a packed batch, concatenated: a cheap way to reach 10^8 alignments.
    Every copy is a distinct set of bundles in distinct memory; only the gene models repeat."""
    if k <= 1:
        return a
    nb, nr = a["b_start"].size, a["r_start"].size
    ns, npair, nj = a["seg_start"].size, a["pair_idx"].size, a["j_start"].size
    width = int(a["b_end"].max()) + gap
    if width * k >= 2 ** 31:
        raise ValueError("replicated coordinates overflow int32")
    out = {}
    sh = (np.arange(k, dtype=np.int64) * width)
    for key in ("b_start", "b_end"):
        out[key] = (a[key].astype(np.int64)[None, :] + sh[:, None]).reshape(-1).astype(np.int32)
    for key in ("r_start", "r_end", "seg_start", "seg_end"):
        out[key] = (a[key].astype(np.int64)[None, :] + sh[:, None]).reshape(-1).astype(np.uint32)
    js = a["j_start"].astype(np.int64)[None, :] + sh[:, None]
    je = a["j_end"].astype(np.int64)[None, :] + sh[:, None]
    sentinel = (a["j_start"] == 0) & (a["j_end"] == 0)
    js[:, sentinel] = 0
    je[:, sentinel] = 0
    out["j_start"] = js.reshape(-1).astype(np.uint32)
    out["j_end"] = je.reshape(-1).astype(np.uint32)
    for key in ("r_len", "r_count", "r_nh", "r_strand", "r_flags", "rjunc_idx", "pair_idx", "pair_cnt", "j_strand",
                "j_guide_match", "j_nreads", "j_nm", "j_mm"):
        out[key] = np.tile(a[key], k)
    for key, n in (("b_read_off", nr), ("b_junc_off", nj), ("r_seg_off", ns), ("r_pair_off", npair)):
        body = a[key].astype(np.int64)[None, :-1] + (np.arange(k, dtype=np.int64) * n)[:, None]
        out[key] = np.concatenate([body.reshape(-1), [k * n]]).astype(np.uint32)
    return out


def concat(parts, gap: int = 100000) -> Dict[str, np.ndarray]:
    """Independent packed batches laid one after the other on the coordinate axis (each shifted past the previous one)."""
    if len(parts) == 1:
        return parts[0]
    out = {k: [] for k in parts[0]}
    shift = 0
    nr = ns = npair = nj = 0
    for i, a in enumerate(parts):
        last = i == len(parts) - 1
        for key in ("b_start", "b_end"):
            out[key].append((a[key].astype(np.int64) + shift).astype(np.int32))
        for key in ("r_start", "r_end", "seg_start", "seg_end"):
            out[key].append((a[key].astype(np.int64) + shift).astype(np.uint32))
        sentinel = (a["j_start"] == 0) & (a["j_end"] == 0)
        for key in ("j_start", "j_end"):
            x = a[key].astype(np.int64) + shift
            x[sentinel] = 0
            out[key].append(x.astype(np.uint32))
        for key in ("r_len", "r_count", "r_nh", "r_strand", "r_flags", "rjunc_idx", "pair_idx", "pair_cnt", "j_strand",
                    "j_guide_match", "j_nreads", "j_nm", "j_mm"):
            out[key].append(a[key])
        for key, n in (("b_read_off", nr), ("b_junc_off", nj), ("r_seg_off", ns), ("r_pair_off", npair)):
            body = a[key].astype(np.int64) + n
            out[key].append((body if last else body[:-1]).astype(np.uint32))
        nr += a["r_start"].size; ns += a["seg_start"].size; npair += a["pair_idx"].size; nj += a["j_start"].size
        shift += int(a["b_end"].max()) + gap - (0 if i else 0)
        if shift >= 2 ** 31 - 2 ** 27:
            raise ValueError("concatenated coordinates overflow int32")
        shift = shift if i == 0 else shift
    # shifts are cumulative: part i starts past the end of part i-1
    return {k: np.concatenate(v) for k, v in out.items()}


def total_input_bytes(a: Dict[str, np.ndarray]) -> int:
    from .stgb import INPUT_FIELDS
    return int(sum(a[k].nbytes for k in INPUT_FIELDS))
