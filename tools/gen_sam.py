#!/usr/bin/env python3
"""Seeded synthetic short-read SAM (+ guide GTF) generator for the reference-side oracle.

Produces coordinate-sorted paired-end spliced alignments in SAM text that the reference binary parses through its
own producer (stringtie.cpp:503-818 -> processRead, rlink.cpp:659), so the packed batches dumped by
oracle/_ref/ref_tool carry exactly what the reference hands to infer_transcripts().  Shape follows SURVEY.md §8d (C2,
scaled down): multi-isoform loci on both strands, log-normal expression, NH in {1,2,3,4,5}, XS only on spliced
reads, duplicates (collapse path, rlink.cpp:757-765), overlapping mates (merge path, rlink.cpp:149-211), a few
long introns (> 100 kb anchor rule, rlink.cpp:16993) and unpaired reads.
"""
import argparse
import random
import sys


def cigar_for(tpos, st, length):
    seg = tpos[st:st + length]
    cig = []
    run = 1
    for a, b in zip(seg, seg[1:]):
        if b == a + 1:
            run += 1
        else:
            cig.append("%dM%dN" % (run, b - a - 1))
            run = 1
    cig.append("%dM" % run)
    return seg[0], "".join(cig), seg[-1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=40)
    ap.add_argument("--pairs", type=int, default=8000)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--readlen", type=int, default=100)
    ap.add_argument("--chroms", type=int, default=2)
    ap.add_argument("--long-intron-frac", type=float, default=0.03)
    ap.add_argument("--nh-frac", type=float, default=0.08, help="fraction of fragments with NH>1")
    ap.add_argument("--jitter-frac", type=float, default=0.0,
                    help="fraction of fragments aligned with one splice site shifted by 1-6 bp and an NM tag above "
                         "mismatchfrac: junctions whose whole support is 'mismatchy' (the higherr pass, rlink.cpp:14538)")
    ap.add_argument("--sam", required=True)
    ap.add_argument("--gtf", default=None)
    a = ap.parse_args()
    rng = random.Random(a.seed)
    RL = a.readlen
    loci = []
    per_chrom = max(1, a.loci // a.chroms)
    for c in range(a.chroms):
        pos0 = 5000
        for g in range(per_chrom):
            nex = rng.randint(1, 12)
            exons = []
            p = pos0
            for e in range(nex):
                L = rng.randint(40, 700)
                exons.append((p, p + L - 1))
                intron = rng.randint(70, 4000)
                if rng.random() < a.long_intron_frac:
                    intron = rng.randint(100001, 140000)
                p += L + intron
            strand = rng.choice("+-")
            isos = [list(range(nex))]
            for k in range(rng.randint(0, 3)):
                if nex > 2:
                    keep = [0] + [i for i in range(1, nex - 1) if rng.random() > 0.3] + [nex - 1]
                    if keep not in isos:
                        isos.append(keep)
            w = [rng.lognormvariate(0, 1) for _ in isos]
            loci.append((c, exons, strand, isos, w, rng.lognormvariate(0, 1.6)))
            # neighbouring loci sometimes closer than runoffdist (bundle merging), sometimes far apart
            pos0 = exons[-1][1] + rng.choice([120, 400, 3000, 20000])
    tw = sum(l[5] for l in loci)
    recs = []
    rid = 0
    gtf = open(a.gtf, "w") if a.gtf else None
    for gi, (c, exons, strand, isos, w, lw) in enumerate(loci):
        chrom = "chr%d" % (c + 1)
        tposs = []
        for ti, iso in enumerate(isos):
            t = []
            for i in iso:
                t += list(range(exons[i][0], exons[i][1] + 1))
            tposs.append(t)
            if gtf:
                ex = [exons[i] for i in iso]
                attr = 'gene_id "G%d"; transcript_id "G%d.%d";' % (gi, gi, ti)
                gtf.write("%s\tsyn\ttranscript\t%d\t%d\t.\t%s\t.\t%s\n" % (chrom, ex[0][0], ex[-1][1], strand, attr))
                for s, e in ex:
                    gtf.write("%s\tsyn\texon\t%d\t%d\t.\t%s\t.\t%s\n" % (chrom, s, e, strand, attr))
        n = max(1, int(a.pairs * lw / tw))
        for _ in range(n):
            ti = rng.choices(range(len(tposs)), w)[0]
            t = tposs[ti]
            nmtag = ""
            if a.jitter_frac and len(isos[ti]) > 1 and rng.random() < a.jitter_frac:
                iso = isos[ti]
                k = rng.randrange(len(iso) - 1)
                d = rng.choice([-6, -5, -4, -3, -2, -1, 1, 2, 3, 4, 5, 6])
                ex = [list(exons[i]) for i in iso]
                if ex[k][1] + d > ex[k][0] + 12 and ex[k + 1][0] + d < ex[k + 1][1] - 12 and ex[k][1] + d < ex[k + 1][0] + d - 20:
                    ex[k][1] += d
                    ex[k + 1][0] += d
                    t = []
                    for s_, e_ in ex:
                        t += list(range(s_, e_ + 1))
                    nmtag = "\tNM:i:%d" % max(3, RL // 25)
            if len(t) < RL + 10:
                continue
            fl = min(len(t), max(RL, int(rng.gauss(2.2 * RL, 0.4 * RL))))
            st = rng.randint(0, len(t) - fl)
            if rng.random() < 0.15:  # snap to a grid so identical alignments occur (collapse path)
                st = (st // 7) * 7
                fl = min(len(t) - st, max(RL, (fl // 20) * 20))
            p1, c1, e1 = cigar_for(t, st, RL)
            p2, c2, e2 = cigar_for(t, st + fl - RL, RL)
            nh = 1
            if rng.random() < a.nh_frac:
                nh = rng.choice([2, 2, 3, 3, 4, 5, 6])
            x1 = ("\tXS:A:" + strand if "N" in c1 else "") + nmtag
            x2 = ("\tXS:A:" + strand if "N" in c2 else "") + nmtag
            name = "r%d" % rid
            rid += 1
            u = rng.random()
            if u < 0.06:  # unpaired single read
                recs.append((c, p1, "%s\t0\t%s\t%d\t60\t%s\t*\t0\t0\t*\t*\tNH:i:%d%s\n" % (name, chrom, p1, c1, nh, x1)))
            else:
                tl = e2 - p1 + 1
                recs.append((c, p1, "%s\t99\t%s\t%d\t60\t%s\t=\t%d\t%d\t*\t*\tNH:i:%d%s\n" % (name, chrom, p1, c1, p2, tl, nh, x1)))
                recs.append((c, p2, "%s\t147\t%s\t%d\t60\t%s\t=\t%d\t%d\t*\t*\tNH:i:%d%s\n" % (name, chrom, p2, c2, p1, -tl, nh, x2)))
    recs.sort(key=lambda x: (x[0], x[1]))
    with open(a.sam, "w") as out:
        out.write("@HD\tVN:1.0\tSO:coordinate\n")
        for c in range(a.chroms):
            out.write("@SQ\tSN:chr%d\tLN:400000000\n" % (c + 1))
        for _, _, r in recs:
            out.write(r)
    if gtf:
        gtf.close()
    print("alignments", len(recs), file=sys.stderr)


if __name__ == "__main__":
    main()
