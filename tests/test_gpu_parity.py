"""Parity of the CUDA path (through the C ABI) with the oracle / the reference's golden vectors.
Bar: bpcov bit-exact (compared as uint32 bit patterns), junction support exactly equal (f64), get_cov exactly equal."""
import numpy as np
import pytest

from stringtie_b200 import synth
from tests import orc
from tests.conftest import load_golden, GOLDEN

pytestmark = pytest.mark.gpu


def gpu_cov_dense(engine, batch, arrays):
    """GPU bpcov re-packed into the oracle's dense layout (lane stride = len)."""
    lens, off = orc.cov_layout(arrays)
    goff, gstride, _ = batch.layout()
    raw = engine.fetch_bpcov_all(batch)
    out = np.empty(int(off[-1]), dtype=np.float32)
    for b in range(lens.size):
        L = int(lens[b])
        for s in range(3):
            out[off[b] + s * L: off[b] + (s + 1) * L] = raw[goff[b] + s * gstride[b]: goff[b] + s * gstride[b] + L]
    return out


A7_KEYS = ("strand", "nreads", "nreads_good", "leftsupport", "rightsupport", "mm", "ej", "good", "good_strand")


def golden_a7(g):
    return {"strand": g["a7_j_strand"], "nreads": g["a7_j_nreads"], "nreads_good": g["a7_j_nreads_good"],
            "leftsupport": g["a7_j_leftsupport"], "rightsupport": g["a7_j_rightsupport"], "mm": g["a7_j_mm"],
            "ej": g["a7_ej"], "good": g["a7_j_good"], "good_strand": g["a7_j_good_strand"]}


def check_batch(engine, arrays, ref_cov=None, ref_support=None, ref_a7=None):
    batch = engine.upload(arrays)
    try:
        engine.run(batch)
        got = gpu_cov_dense(engine, batch, arrays)
        good, left, right = engine.fetch_junction_support(batch)
        if ref_cov is None:
            ref_cov, _, g, l, r = orc.run(arrays)
            ref_support = (g, l, r)
        bad = np.nonzero(got.view(np.uint32) != ref_cov.view(np.uint32))[0]
        assert bad.size == 0, f"{bad.size} bpcov entries differ, first at {bad[:5]}: {got[bad[:5]]} vs {ref_cov[bad[:5]]}"
        assert np.array_equal(good, ref_support[0])
        assert np.array_equal(left, ref_support[1])
        assert np.array_equal(right, ref_support[2])
        # row a7: junction pass of build_graphs + good_junc, against the oracle or the reference's own dump
        jp = engine.fetch_junction_pass(batch)
        if ref_a7 is None:
            lens_, off_ = orc.cov_layout(arrays)
            ref_a7 = orc.junction_pass(arrays, ref_cov, off_, *ref_support)
        for k in A7_KEYS:
            bad = np.nonzero(jp[k] != ref_a7[k])[0]
            assert bad.size == 0, f"a7 {k}: {bad.size} rows differ, first {bad[:5]}: {jp[k][bad[:5]]} vs {ref_a7[k][bad[:5]]}"
        # second run on the same device batch: identical (idempotent, no leftover state in scratch)
        engine.run(batch)
        got2 = gpu_cov_dense(engine, batch, arrays)
        assert np.array_equal(got.view(np.uint32), got2.view(np.uint32))
        g2, l2, r2 = engine.fetch_junction_support(batch)
        assert np.array_equal(good, g2) and np.array_equal(left, l2) and np.array_equal(right, r2)
        return batch
    except Exception:
        batch.free()
        raise


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_reference_vectors(engine, name):
    g = load_golden(name)
    b = check_batch(engine, g, g["s1_cov"], (g["s1_j_nreads_good"], g["s1_j_leftsupport"], g["s1_j_rightsupport"]),
                    golden_a7(g))
    b.free()


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_adversarial_vs_oracle(engine, seed):
    a = synth.make_adversarial(seed=seed, n_bundles=14, neg_links=(seed % 2 == 0))
    check_batch(engine, a).free()


@pytest.mark.parametrize("nb,rpb,seed", [(50, 600, 1), (300, 1700, 2), (20, 20000, 3)])
def test_synthetic_c2_shape_vs_oracle(engine, nb, rpb, seed):
    a = synth.make_batch(n_bundles=nb, reads_per_bundle=rpb, seed=seed)
    check_batch(engine, a).free()


def test_single_bundle_drop_in(engine):
    """The reference-shaped call: one bundle in, what infer_transcripts writes back out."""
    g = load_golden("short_s11")
    lens, off = orc.cov_layout(g)
    for b in (0, 5, int(np.argmax(lens))):
        cov, good, left, right = engine.infer_transcripts(g, b)
        L = int(lens[b])
        want = g["s1_cov"][off[b]:off[b + 1]].reshape(3, L)
        assert np.array_equal(cov.view(np.uint32), want.view(np.uint32))
        j0, j1 = int(g["b_junc_off"][b]), int(g["b_junc_off"][b + 1])
        assert np.array_equal(good, g["s1_j_nreads_good"][j0:j1])
        assert np.array_equal(left, g["s1_j_leftsupport"][j0:j1])
        assert np.array_equal(right, g["s1_j_rightsupport"][j0:j1])


def test_get_cov_queries(engine):
    g = load_golden("short_s11")
    lens, off = orc.cov_layout(g)
    batch = engine.upload(g)
    engine.run(batch)
    rng = np.random.default_rng(5)
    nq = 4000
    bun = rng.integers(0, lens.size, nq)
    lane = rng.integers(0, 3, nq)
    start = (rng.random(nq) * (lens[bun] - 3)).astype(np.int64)
    end = start + (rng.random(nq) * (lens[bun] - 2 - start)).astype(np.int64)
    for sign in (False, True):
        got = engine.query_cov(batch, bun, lane, start, end, sign=sign)
        for i in range(nq):
            b = int(bun[i])
            want = orc.get_cov(g["s1_cov"][off[b]:off[b + 1]], int(lens[b]), int(lane[i]), int(start[i]), int(end[i]), sign)
            assert got[i] == want, (i, sign, got[i], want)
    tot = engine.fetch_cov_totals(batch)
    for b in range(lens.size):
        for s in range(3):
            assert tot[b, s] == orc.get_cov(g["s1_cov"][off[b]:off[b + 1]], int(lens[b]), s, 0, int(lens[b]) - 2)
    batch.free()


def test_empty_and_degenerate_batches(engine):
    a = synth.make_adversarial(seed=9, n_bundles=4)
    # a batch whose bundles have no reads at all
    from stringtie_b200.stgb import INPUT_FIELDS
    e = {k: np.zeros(0, dtype=dt) for k, dt in INPUT_FIELDS.items()}
    e["b_start"] = np.array([100, 5000], dtype=np.int32)
    e["b_end"] = np.array([1500, 5010], dtype=np.int32)
    e["b_read_off"] = np.zeros(3, dtype=np.uint32)
    e["b_junc_off"] = np.zeros(3, dtype=np.uint32)
    e["r_seg_off"] = np.zeros(1, dtype=np.uint32)
    e["r_pair_off"] = np.zeros(1, dtype=np.uint32)
    batch = engine.upload(e)
    engine.run(batch)
    assert not engine.fetch_bpcov_all(batch).any()
    batch.free()
    check_batch(engine, a).free()


def test_large_property_checks(engine):
    """~3M alignments: full comparison against the oracle plus size-independent properties of bpcov:
    (1) within every BSIZE block the stored prefix sums are non-decreasing (clamped coverage is >= 0);
    (2) lane 1 dominates lanes 0 and 2 (every stranded deposit is mirrored on lane 1);
    (3) total coverage == sum over deposited intervals of w * length (float64, tolerance 1e-4 relative)."""
    a = synth.make_batch(n_bundles=1800, reads_per_bundle=1700, seed=11)
    batch = check_batch(engine, a)
    tot = engine.fetch_cov_totals(batch)
    assert (tot[:, 1] + 1e-3 >= tot[:, 0]).all() and (tot[:, 1] + 1e-3 >= tot[:, 2]).all()
    raw = engine.fetch_bpcov_all(batch)
    goff, gstride, _ = batch.layout()
    lens, _ = orc.cov_layout(a)
    for b in np.random.default_rng(0).integers(0, lens.size, 40):
        L = int(lens[b])
        for s in range(3):
            c = raw[goff[b] + s * gstride[b]: goff[b] + s * gstride[b] + L].astype(np.float64)
            d = np.diff(c)
            d[9999::10000] = 0  # block restarts
            assert (d >= 0).all()
    batch.free()


def test_c2_full_size(engine):
    """BASELINE.json's full size (the bench workload: ~10^8 alignments, 60 000 bundles), every stage against the
    oracle: per-bundle coverage totals (get_cov over the whole bundle: any wrong bit in the blocked prefix sums moves
    them), junction support and the junction pass on every row, and the full bpcov bit pattern of the deepest
    bundles plus a random sample."""
    import bench
    base, a = bench.build_workload(bench.REPLICAS, bench.BASE_BUNDLES, 0)
    batch = engine.upload(a)
    engine.run(batch)
    tot = engine.fetch_cov_totals(batch)
    good, left, right = engine.fetch_junction_support(batch)
    jp = engine.fetch_junction_pass(batch)
    cov, off, og, ol, orr = orc.run(a, threads=0)
    assert np.array_equal(good, og) and np.array_equal(left, ol) and np.array_equal(right, orr)
    lens, _ = orc.cov_layout(a)
    nb = lens.size
    reads = np.diff(a["b_read_off"].astype(np.int64))
    sample = np.unique(np.concatenate([np.argsort(reads)[-24:], np.argsort(lens)[-8:],
                                       np.random.default_rng(1).integers(0, nb, 300)]))
    for b in sample:
        L = int(lens[b])
        got = engine.fetch_bpcov(batch, int(b), L).reshape(3, L)
        want = cov[off[b]:off[b + 1]].reshape(3, L)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"bpcov of bundle {b} differs"
    # totals of ALL bundles, recomputed from the oracle's bpcov with the oracle's get_cov on a sample, and cross-lane
    for b in sample:
        for s in range(3):
            assert tot[b, s] == orc.get_cov(cov[off[b]:off[b + 1]], int(lens[b]), s, 0, int(lens[b]) - 2)
    # every bundle: the stored last prefix of every BSIZE block summed in f64 == the GPU total (vectorised get_cov)
    for s in (1,):
        want_all = np.zeros(nb)
        for b in range(nb):
            L = int(lens[b])
            lane = cov[off[b] + s * L: off[b] + (s + 1) * L]
            m = (L - 1) // 10000
            acc = float(lane[9999:m * 10000:10000].astype(np.float64).sum()) if m else 0.0
            acc += float(np.float32(lane[L - 1]) - np.float32(lane[0]))
            want_all[b] = max(acc, 0.0)
        assert np.array_equal(tot[:, s], want_all)
    ref_a7 = orc.junction_pass(a, cov, off, og, ol, orr)
    for k in A7_KEYS:
        assert np.array_equal(jp[k], ref_a7[k]), f"a7 {k} differs at full size"
    batch.free()
    engine.trim()


def test_malformed_batches_are_rejected_not_executed(engine):
    """Indices that point outside their bundle must come back as STG_EINVAL from stg_run (the kernels skip them instead
    of reading out of bounds), and the context must stay usable."""
    from stringtie_b200.engine import StgError
    good = synth.make_batch(n_bundles=6, reads_per_bundle=300, seed=21)
    for field, value, msg in (("pair_idx", 10 ** 6, "pair_idx"), ("rjunc_idx", 10 ** 6, "rjunc_idx")):
        a = {k: v.copy() for k, v in good.items()}
        assert a[field].size > 10
        a[field][a[field].size // 2] = value
        batch = engine.upload(a)
        with pytest.raises(StgError, match=msg):
            engine.run(batch)
        batch.free()
    a = {k: v.copy() for k, v in good.items()}
    a["r_seg_off"][5] = a["r_seg_off"][4] - 1 if a["r_seg_off"][4] > 0 else a["r_seg_off"][6] + 7
    with pytest.raises(StgError, match="segments"):     # refused at upload: nothing reaches a kernel
        engine.upload(a)
    # a read without segments followed by a spliced read: r_seg_off[n] - n would wrap in the kernels
    a = {k: v.copy() for k, v in good.items()}
    a["r_seg_off"][1] = a["r_seg_off"][0]
    with pytest.raises(StgError, match="segments"):
        engine.upload(a)
    # reads without a bundle
    a = {k: v.copy() for k, v in good.items()}
    a["b_start"] = a["b_start"][:0]; a["b_end"] = a["b_end"][:0]
    a["b_read_off"] = a["b_read_off"][:1]; a["b_junc_off"] = a["b_junc_off"][:1]
    with pytest.raises(StgError, match="without a bundle"):
        engine.upload(a)
    check_batch(engine, good).free()   # the context still works


def _regions_from_oracle(a, cov, off, good, left, right):
    """Bundlenodes of the oracle (rows a8 + a9) and their halves: the kind of regions create_graph hands to find_all_trims."""
    bun, sno, st, en = [], [], [], []
    for b, g in enumerate(orc.groups(a, cov, off, good, left, right)):
        for s in range(3):
            for k in range(int(g["bn_off"][s]), int(g["bn_off"][s + 1])):
                x, z = int(g["bn_start"][k]), int(g["bn_end"][k])
                m = x + (z - x) // 2
                for lo, hi in ((x, z), (x, m), (m + 1, z)):
                    if hi > lo:
                        bun.append(b); sno.append(2 if s == 2 else 0); st.append(lo); en.append(hi)
    return np.array(bun, np.int32), np.array(sno, np.int8), np.array(st, np.uint32), np.array(en, np.uint32)


@pytest.mark.parametrize("name", GOLDEN)
def test_find_trims_golden(engine, name):
    """find_all_trims on the GPU against the reference's own results (dumped with the fixtures)."""
    g = load_golden(name)
    from stringtie_b200.stgb import input_view
    batch = engine.upload(input_view(g))
    engine.run(batch)
    off, pos, ab, st = engine.find_trims(batch, g["ft_bundle"], g["ft_sno"], g["ft_start"], g["ft_end"])
    assert np.array_equal(off, g["ft_off"]) and np.array_equal(pos, g["ft_pos"]) and np.array_equal(st, g["ft_isstart"])
    assert np.array_equal(ab.view(np.uint32), g["ft_abund"].view(np.uint32))
    batch.free()


@pytest.mark.parametrize("kind,seed", [("synth", 3), ("deep", 4), ("adv", 1), ("adv", 3)])
def test_find_trims_vs_oracle(engine, kind, seed):
    a = (synth.make_batch(n_bundles=60, reads_per_bundle=3000, seed=seed) if kind == "synth" else
         synth.make_batch(n_bundles=12, reads_per_bundle=40000, seed=seed) if kind == "deep" else
         synth.make_adversarial(seed=seed, n_bundles=14))
    cov, off, good, left, right = orc.run(a)
    bun, sno, st, en = _regions_from_oracle(a, cov, off, good, left, right)
    assert bun.size > 50
    batch = engine.upload(a)
    engine.run(batch)
    got = engine.find_trims(batch, bun, sno, st, en)
    want = orc.find_trims(a, cov, off, bun, sno, st, en)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[3], want[3])
    assert np.array_equal(got[2].view(np.uint32), want[2].view(np.uint32))
    batch.free()


# ---- rows a8 + a9: read repair, groups, colours, bundles / bundlenodes (stg_build_groups) ----------------------
def _gpu_groups(engine, a):
    from tests import a9_util
    batch = engine.upload(a)
    try:
        engine.run(batch)
        engine.build_groups(batch)
        return a9_util.split_engine_groups(engine.fetch_groups(batch), a)
    finally:
        batch.free()


@pytest.mark.parametrize("name", GOLDEN)
def test_groups_golden(engine, name):
    """State after the read loop, the colour pass and the bundle creation against the dumps taken INSIDE the reference's
    build_graphs (hooks at rlink.cpp:15194 and :15433), committed with the fixtures."""
    from tests import a9_util
    from stringtie_b200.stgb import input_view
    g = load_golden(name)
    a = input_view(g)
    a9_util.assert_same(_gpu_groups(engine, a), a9_util.split_reference_dump(g, a))


@pytest.mark.parametrize("kind,seed", [("synth", 3), ("synth", 7), ("deep", 4), ("adv", 1), ("adv", 2), ("adv", 3), ("adv", 5)])
def test_groups_vs_oracle(engine, kind, seed):
    from tests import a9_util
    a = (synth.make_batch(n_bundles=60, reads_per_bundle=3000, seed=seed) if kind == "synth" else
         synth.make_batch(n_bundles=8, reads_per_bundle=30000, seed=seed) if kind == "deep" else
         synth.make_adversarial(seed=seed, n_bundles=14))
    cov, off, good, left, right = orc.run(a)
    a9_util.assert_same(_gpu_groups(engine, a), orc.groups(a, cov, off, good, left, right))


@pytest.fixture(scope="module")
def cluster_engine():
    """A context whose oversized-bundle threshold (row x1) is lowered, so that ordinary test bundles are walked by the
    thread-block-cluster kernel: its window, its grow-inside-the-run dependencies, its logged ordered sums."""
    import os
    from stringtie_b200.engine import Engine
    old = os.environ.get("STGPU_CLUSTER_READS")
    os.environ["STGPU_CLUSTER_READS"] = "500"
    try:
        e = Engine(0)
    finally:
        if old is None:
            del os.environ["STGPU_CLUSTER_READS"]
        else:
            os.environ["STGPU_CLUSTER_READS"] = old
    yield e
    e.close()


@pytest.mark.parametrize("kind,seed", [("synth", 3), ("deep", 4), ("deep", 9), ("adv", 1), ("adv", 2), ("adv", 5), ("sparse", 21)])
def test_groups_cluster_kernel_vs_oracle(cluster_engine, kind, seed):
    from tests import a9_util
    a = (synth.make_batch(n_bundles=30, reads_per_bundle=3000, seed=seed) if kind == "synth" else
         synth.make_batch(n_bundles=8, reads_per_bundle=30000, seed=seed) if kind == "deep" else
         synth.make_sparse(n_reads=40000, seed=seed, n_bundles=2) if kind == "sparse" else
         synth.make_adversarial(seed=seed, n_bundles=14, max_reads=3000))
    cov, off, good, left, right = orc.run(a)
    a9_util.assert_same(_gpu_groups(cluster_engine, a), orc.groups(a, cov, off, good, left, right))


@pytest.mark.parametrize("name", GOLDEN)
def test_groups_cluster_kernel_golden(cluster_engine, name):
    from tests import a9_util
    from stringtie_b200.stgb import input_view
    g = load_golden(name)
    a = input_view(g)
    a9_util.assert_same(_gpu_groups(cluster_engine, a), a9_util.split_reference_dump(g, a))


@pytest.mark.parametrize("seed", [38, 43])
def test_groups_appended_junctions(engine, seed):
    """Seeds on which the repair appends junction rows and the table is re-sorted with the reference's quicksort (the
    oracle is pinned on exactly these inputs against the live reference, tests/test_oracle_vs_reference.py)."""
    from tests import a9_util
    a = synth.make_adversarial(seed=seed, n_bundles=10, max_reads=250)
    rng = np.random.default_rng(seed)
    a["j_nm"] = np.where(rng.random(a["j_nm"].size) < 0.5, a["j_nreads"], a["j_nm"])
    cov, off, good, left, right = orc.run(a)
    want = orc.groups(a, cov, off, good, left, right)
    assert sum(w["j_start"].size for w in want) > a["j_start"].size
    a9_util.assert_same(_gpu_groups(engine, a), want)


def test_groups_bundledist_zero_and_empty():
    """bundledist = 0 (no merging of neighbouring groups) and a batch with empty bundles."""
    from tests import a9_util
    from stringtie_b200.engine import Engine, default_params
    p = default_params()
    p.bundledist = 0
    e = Engine(0, p)
    try:
        a = synth.make_adversarial(seed=9, n_bundles=12)
        cov, off, good, left, right = orc.run(a)
        a9_util.assert_same(_gpu_groups(e, a), orc.groups(a, cov, off, good, left, right, bundledist=0))
    finally:
        e.close()


def test_groups_refuses_unsupported_modes():
    from stringtie_b200.engine import Engine, StgError, default_params
    p = default_params()
    p.guided = 1   # guide boundaries change the merging of groups (rlink.cpp:15134-15164): not built
    e = Engine(0, p)
    try:
        batch = e.upload(synth.make_batch(n_bundles=3, reads_per_bundle=50, seed=1))
        e.run(batch)
        with pytest.raises(StgError):
            e.build_groups(batch)
        batch.free()
    finally:
        e.close()


@pytest.mark.parametrize("n_reads", [3000, 17000])   # walked by the 2-warp kernel / by the 4-warp kernel of deep bundles
def test_groups_many_groups_leave_shared_memory(engine, n_reads):
    """Bundles whose group arrays outgrow the shared-memory slots of the walk kernel (sparse reads over long spans:
    thousands of separate groups per strand lane) carry on in global memory."""
    from tests import a9_util
    a = synth.make_sparse(n_reads=n_reads, seed=21, n_bundles=3 if n_reads < 10000 else 2)
    cov, off, good, left, right = orc.run(a)
    want = orc.groups(a, cov, off, good, left, right)
    assert min(w["g_start"].size for w in want) > 600
    a9_util.assert_same(_gpu_groups(engine, a), want)


def test_groups_c2_base_sample(engine):
    """The seeded base of the bench workload (6000 loci, ~10^7 alignments): the three deepest bundles and a random
    sample of 150 others, every field equal to the oracle; the whole base is walked on the GPU."""
    import bench
    from stringtie_b200 import shard
    from tests import a9_util
    base, _ = bench.build_workload(1, bench.BASE_BUNDLES, 0)
    nreads = np.diff(base["b_read_off"].astype(np.int64))
    rng = np.random.default_rng(7)
    pick = np.unique(np.concatenate([rng.choice(nreads.size, size=150, replace=False), np.argsort(nreads)[-3:]]))
    batch = engine.upload(base)
    try:
        engine.run(batch)
        engine.build_groups(batch)
        G = engine.fetch_groups(batch)
        N = engine.neutral_predictions(batch)
    finally:
        batch.free()
    got_all = a9_util.split_engine_groups(G, base)
    sub = shard.take_bundles(base, pick)
    cov, off, good, left, right = orc.run(sub, threads=0)
    want = orc.groups(sub, cov, off, good, left, right)
    a9_util.assert_same([got_all[int(b)] for b in pick], want)
    # row a19 on the same sample: the neutral bundles' predictions of every picked bundle
    wantn = orc.neutral_preds(sub, cov, off, good, left, right)
    noff = N["off"].astype(np.int64)
    assert noff[-1] > 100
    for k, b in enumerate(pick):
        p0, p1 = int(noff[b]), int(noff[b + 1])
        for key in ("start", "end", "cov", "tlen", "geneno"):
            assert np.array_equal(N[key][p0:p1], wantn[k][key].astype(N[key].dtype)), (int(b), key)
    # size-independent properties over ALL bundles: live groups of a lane are disjoint and sorted along the chain
    go = G["GROUP_OFF"].astype(np.int64)
    for b in rng.choice(nreads.size, size=400, replace=False):
        g0, g1 = int(go[b]), int(go[b + 1])
        st, en, nx = G["G_START"][g0:g1].astype(np.int64), G["G_END"][g0:g1].astype(np.int64), G["G_NEXT"][g0:g1]
        for s in range(3):
            g = int(G["STARTGROUP"][3 * b + s])
            last_end = -1
            while g != -1:
                assert G["G_ALIVE"][g0 + g] == 1 and st[g] > last_end and en[g] >= st[g]
                last_end = int(en[g]) + 50   # neighbours closer than bundledist were merged
                g = int(nx[g])


def test_groups_call_order_is_checked(engine):
    from stringtie_b200.engine import StgError
    a = synth.make_batch(n_bundles=4, reads_per_bundle=200, seed=2)
    batch = engine.upload(a)
    try:
        with pytest.raises(StgError):          # the junction pass has not run yet
            engine.build_groups(batch)
        with pytest.raises(StgError):          # nothing to fetch either
            engine.fetch_groups_array(batch, "G_START")
        engine.run(batch)
        engine.build_groups(batch)
        with pytest.raises(StgError):          # once per batch
            engine.build_groups(batch)
        assert engine.fetch_groups_array(batch, "GROUP_OFF").size == 5
    finally:
        batch.free()


def test_groups_empty_bundles_and_reads_without_junctions(engine):
    """Bundles without reads, bundles without junction rows, a batch that is all of that."""
    from tests import a9_util
    from stringtie_b200 import shard
    a = synth.make_adversarial(seed=4, n_bundles=10)         # bundle 3 has no reads
    cov, off, good, left, right = orc.run(a)
    a9_util.assert_same(_gpu_groups(engine, a), orc.groups(a, cov, off, good, left, right))
    only_empty = shard.take_bundles(a, np.array([3]))
    got = _gpu_groups(engine, only_empty)
    assert got[0]["g_start"].size == 0 and got[0]["rg"].size == 0 and list(got[0]["startgroup"]) == [-1, -1, -1]
    b = synth.make_sparse(n_reads=50, seed=1, n_bundles=2)  # no junction rows at all
    cov, off, good, left, right = orc.run(b)
    a9_util.assert_same(_gpu_groups(engine, b), orc.groups(b, cov, off, good, left, right))


# ---- row a19 (de novo part): single-exon predictions of the neutral bundles (stg_neutral_predictions) -------------
def _gpu_neutral(engine, a):
    batch = engine.upload(a)
    try:
        engine.run(batch)
        engine.build_groups(batch)
        return engine.neutral_predictions(batch)
    finally:
        batch.free()


@pytest.mark.parametrize("name", GOLDEN)
def test_neutral_predictions_golden(engine, name):
    """Against the prediction list dumped INSIDE the reference's build_graphs right after its neutral-bundle section
    (hook before rlink.cpp:15637): coordinates, length, gene number equal, coverage equal as f64."""
    from stringtie_b200.stgb import input_view
    g = load_golden(name)
    got = _gpu_neutral(engine, input_view(g))
    assert np.array_equal(got["off"], g["a19_pred_off"])
    assert np.array_equal(got["start"], g["a19_start"]) and np.array_equal(got["end"], g["a19_end"])
    assert np.array_equal(got["tlen"], g["a19_tlen"]) and np.array_equal(got["geneno"], g["a19_geneno"])
    assert np.array_equal(got["cov"], g["a19_cov"])
    assert np.array_equal(got["cov"].astype(np.float32), g["a19_exoncov"])


@pytest.mark.parametrize("kind,seed", [("wide", 8), ("wide", 9), ("synth", 3), ("adv", 2)])
def test_neutral_predictions_vs_oracle(engine, kind, seed):
    a = (synth.make_batch(n_bundles=400, reads_per_bundle=500, seed=seed) if kind == "wide" else
         synth.make_batch(n_bundles=60, reads_per_bundle=3000, seed=seed) if kind == "synth" else
         synth.make_adversarial(seed=seed, n_bundles=14))
    cov, off, good, left, right = orc.run(a)
    want = orc.neutral_preds(a, cov, off, good, left, right)
    got = _gpu_neutral(engine, a)
    woff = np.concatenate([[0], np.cumsum([w["start"].size for w in want])]).astype(np.uint32)
    assert np.array_equal(got["off"], woff)
    if kind == "wide":
        assert woff[-1] > 20
    for k in ("start", "end", "cov", "tlen", "geneno"):
        w = np.concatenate([x[k] for x in want]) if want else np.zeros(0)
        assert np.array_equal(got[k], w.astype(got[k].dtype)), k


def test_neutral_predictions_need_groups(engine):
    from stringtie_b200.engine import StgError
    batch = engine.upload(synth.make_batch(n_bundles=3, reads_per_bundle=100, seed=1))
    try:
        engine.run(batch)
        with pytest.raises(StgError):
            engine.neutral_predictions(batch)
    finally:
        batch.free()
