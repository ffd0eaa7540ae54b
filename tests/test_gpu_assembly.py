"""Rows a10-a15 on the GPU (stg_assemble) against the reference itself: golden fixtures dumped by ref_tool and live
ref_hot --full runs on seeded batches.  Bar: every integer equal, every float (node coverage, prediction and exon
coverage) bit-equal."""
import numpy as np
import pytest

from stringtie_b200 import synth
from tests import asm_util
from tests.conftest import load_golden, GOLDEN

pytestmark = pytest.mark.gpu
need_ref = pytest.mark.skipif(not asm_util.have_ref_hot(), reason="oracle/_ref/ref_hot not built")


def _no_unitigs(a):
    a = dict(a)
    a["r_flags"] = (a["r_flags"] & ~np.uint8(1)).astype(np.uint8)
    return a


@pytest.mark.parametrize("name", GOLDEN)
def test_assembly_golden(engine, name):
    g = load_golden(name)
    neu, pr, stage = asm_util.gpu_predictions(engine, g)
    asm_util.assert_same_graph_stage(stage, g)
    asm_util.assert_same_predictions(neu, pr, g)


@need_ref
@pytest.mark.parametrize("nb,rpb,seed", [(64, 1700.0, 7), (300, 1700.0, 8), (40, 12000.0, 9), (400, 300.0, 10)])
def test_assembly_synthetic_vs_reference(engine, nb, rpb, seed):
    a = synth.make_batch(n_bundles=nb, reads_per_bundle=rpb, seed=seed)
    ref = asm_util.ref_full(a)
    neu, pr, stage = asm_util.gpu_predictions(engine, a)
    asm_util.assert_same_graph_stage(stage, ref)
    asm_util.assert_same_predictions(neu, pr, ref)


@need_ref
@pytest.mark.parametrize("seed", [4, 5, 6, 8])
def test_assembly_adversarial_vs_reference(engine, seed):
    a = _no_unitigs(synth.make_adversarial(seed=seed, n_bundles=40, max_reads=600))
    ref = asm_util.ref_full(a)
    try:
        neu, pr, stage = asm_util.gpu_predictions(engine, a)
    except Exception as e:   # a read spanning more graph nodes than the kernels allow is refused, not mis-assembled
        assert "spans more graph nodes" in str(e), e
        return
    asm_util.assert_same_graph_stage(stage, ref)
    asm_util.assert_same_predictions(neu, pr, ref)


def test_assembly_refuses_unitigs_and_call_order(engine):
    from stringtie_b200.stgb import input_view
    a = input_view(synth.make_adversarial(seed=3, n_bundles=6, max_reads=200))
    a["r_flags"][:] |= 1
    batch = engine.upload(a)
    with pytest.raises(Exception):
        engine.assemble(batch)        # groups not built
    engine.run(batch)
    engine.build_groups(batch)
    engine.neutral_predictions(batch)
    with pytest.raises(Exception) as ei:
        engine.assemble(batch)
    assert "unitig" in str(ei.value)
    batch.free()


def test_assembly_idempotent_across_batches(engine):
    """The same input assembled twice (pool blocks reused in between) gives identical predictions."""
    a = synth.make_batch(n_bundles=48, reads_per_bundle=900.0, seed=21)
    _, p1, s1 = asm_util.gpu_predictions(engine, a)
    _, p2, s2 = asm_util.gpu_predictions(engine, a)
    for k in p1:
        assert np.array_equal(p1[k].view(np.uint8), p2[k].view(np.uint8)), k
    assert np.array_equal(s1["N_COV"].view(np.uint32), s2["N_COV"].view(np.uint32))


@need_ref
def test_assembly_deepest_bundles_of_the_bench_workload(engine):
    """The deepest bundles of one part of the bench workload (hundreds of thousands of reads in one locus: the CTA paths of
    the commit stage, the wide kernels of the group stage) plus a random sample, against the reference."""
    import bench
    from stringtie_b200 import shard
    base, _ = bench.build_workload(1, bench.PART_BUNDLES, 0)
    nr = np.diff(base["b_read_off"].astype(np.int64))
    order = np.argsort(-nr)
    rng = np.random.default_rng(5)
    pick = np.unique(np.concatenate([order[:4], rng.choice(nr.size, 60, replace=False)]))
    a = shard.take_bundles(base, pick)
    assert int(np.diff(a["b_read_off"].astype(np.int64)).max()) >= 100000
    ref = asm_util.ref_full(a)
    neu, pr, stage = asm_util.gpu_predictions(engine, a)
    asm_util.assert_same_graph_stage(stage, ref)
    asm_util.assert_same_predictions(neu, pr, ref)


def test_one_bundle_drop_in_returns_the_reference_prediction_list(engine):
    """stg_infer_transcripts(assemble = 1), the call with the shape of infer_transcripts(BundleData*): BundleData::pred,
    bpcov, num_fragments and frag_len of single bundles against the reference's dumps."""
    g = load_golden("short_s13_jitter")
    from stringtie_b200.stgb import input_view
    a = input_view(g)
    for b in (0, 3, 7, 11, int(g["b_start"].size) - 1):
        cov, pred, nfrag, flen = engine.infer_transcripts_full(a, b)
        r0, r1 = int(g["f_pred_off"][b]), int(g["f_pred_off"][b + 1])
        assert np.array_equal(pred["start"], g["p_start"][r0:r1]) and np.array_equal(pred["end"], g["p_end"][r0:r1])
        assert np.array_equal(pred["cov"].view(np.uint64), g["p_cov"][r0:r1].view(np.uint64))
        assert np.array_equal(pred["tlen"], g["p_tlen"][r0:r1]) and np.array_equal(pred["geneno"], g["p_geneno"][r0:r1])
        assert np.array_equal(pred["strand"], g["p_strand"][r0:r1])
        e0, e1 = int(g["p_exon_off"][r0]), int(g["p_exon_off"][r1])
        assert np.array_equal(pred["e_start"], g["pe_start"][e0:e1]) and np.array_equal(pred["e_end"], g["pe_end"][e0:e1])
        assert np.array_equal(pred["e_cov"].view(np.uint32), g["pe_cov"][e0:e1].view(np.uint32))
        assert np.array_equal(pred["exon_off"] - pred["exon_off"][0], g["p_exon_off"][r0:r1 + 1] - g["p_exon_off"][r0])
        assert nfrag == g["f_num_fragments"][b] and flen == g["f_frag_len"][b]
        c0, c1 = int(g["s1_cov_off"][b]), int(g["s1_cov_off"][b + 1])
        assert np.array_equal(cov.reshape(-1).view(np.uint32), g["s1_cov"][c0:c1].view(np.uint32))


def test_batched_drop_in_equals_one_bundle_calls(engine):
    """stg_infer_transcripts_batch (several bundles of the host's worker threads in one device batch) returns for every
    bundle exactly what the one-bundle call returns: BundleData::pred, bpcov, num_fragments, frag_len."""
    g = load_golden("short_s13_jitter")
    from stringtie_b200.stgb import input_view
    a = input_view(g)
    pick = [0, 3, 4, 7, 11, int(g["b_start"].size) - 1]
    got = engine.infer_transcripts_batch(a, pick)
    assert sum(p["start"].size for _, p, _, _ in got) > 10
    for (cov, pred, nfrag, flen), b in zip(got, pick):
        cov1, pred1, nfrag1, flen1 = engine.infer_transcripts_full(a, b)
        assert np.array_equal(cov.view(np.uint32), cov1.view(np.uint32))
        assert nfrag == nfrag1 and flen == flen1
        for k in pred1:
            assert np.array_equal(pred[k].view(np.uint8), pred1[k].view(np.uint8)), (b, k)


def test_packed_bpcov_unpacks_to_the_full_copy(engine):
    """stg_pack_bpcov / stg_fetch_bpcov_packed_begin / stg_unpack_bpcov (the end-to-end path of a batch consumer: sub-tile lanes
    that repeat one value travel as that value) against the plain copy of all of bpcov, bit for bit; bundles longer than a
    BSIZE block, ragged last sub-tiles, empty bundles."""
    for a in (synth.make_batch(n_bundles=300, reads_per_bundle=900, seed=21), synth.make_adversarial(seed=5, n_bundles=30, max_reads=400)):
        batch = engine.upload(a)
        engine.run(batch)
        full = engine.fetch_bpcov_all(batch)
        nt, nraw = engine.pack_bpcov(batch)
        assert 0 < nraw < 3 * nt
        desc = np.zeros(3 * nt, np.uint32); val = np.zeros(3 * nt, np.float32); raw = np.zeros(max(1, nraw) * 512, np.float32)
        engine.fetch_bpcov_packed_begin(batch, desc.ctypes.data, val.ctypes.data, raw.ctypes.data)
        engine.fetch_wait(batch)
        got = np.full(full.size, np.float32(-1.0))
        nb = int(a["b_start"].size)
        engine.unpack_bpcov(batch, 0, nb // 2, desc.ctypes.data, val.ctypes.data, raw.ctypes.data, got.ctypes.data)
        engine.unpack_bpcov(batch, nb // 2, nb, desc.ctypes.data, val.ctypes.data, raw.ctypes.data, got.ctypes.data)
        off, stride, tot = batch.layout()
        lens = a["b_end"].astype(np.int64) - a["b_start"].astype(np.int64) + 3
        for b in range(nb):
            for s in range(3):
                o = int(off[b]) + s * int(stride[b])
                assert np.array_equal(got[o:o + lens[b]].view(np.uint32), full[o:o + lens[b]].view(np.uint32)), (b, s)
        # the per-bundle form a host adaptor's worker calls, and a worker's share of the batch through one set of working arrays
        for b in (0, nb // 3, nb - 1):
            one = engine.unpack_bpcov_bundle(batch, b, int(lens[b]), desc, val, raw)
            for s in range(3):
                o = int(off[b]) + s * int(stride[b])
                assert np.array_equal(one[s].view(np.uint32), full[o:o + lens[b]].view(np.uint32))
        work = np.zeros(3 * int(lens.max()), np.float32)
        chk = engine.unpack_bpcov_stream(batch, 0, nb, desc.ctypes.data, val.ctypes.data, raw.ctypes.data, work)
        want = sum(float(full[int(off[b]) + int(stride[b]) + int(lens[b]) - 1]) for b in range(nb))
        assert chk == want
        batch.free()


def test_release_device_keeps_the_handle_for_unpacking(engine):
    """stg_release_device: the device memory of a batch goes back to the pool (no cudaFree), the handle still unpacks bpcov, and the
    stages refuse to run on it."""
    a = synth.make_batch(n_bundles=40, reads_per_bundle=500, seed=3)
    batch = engine.upload(a)
    engine.run(batch)
    full = engine.fetch_bpcov_all(batch)
    nt, nraw = engine.pack_bpcov(batch)
    desc = np.zeros(3 * nt, np.uint32); val = np.zeros(3 * nt, np.float32); raw = np.zeros(max(1, nraw) * 512, np.float32)
    engine.fetch_bpcov_packed_begin(batch, desc.ctypes.data, val.ctypes.data, raw.ctypes.data)
    engine.fetch_wait(batch)
    before = engine.pool_stats()
    engine.release_device(batch)
    after = engine.pool_stats()
    assert after["pool_bytes"] > before["pool_bytes"] and after["cudaFree_calls"] == before["cudaFree_calls"]
    off, stride, tot = batch.layout()
    lens = a["b_end"].astype(np.int64) - a["b_start"].astype(np.int64) + 3
    one = engine.unpack_bpcov_bundle(batch, 7, int(lens[7]), desc, val, raw)
    for s in range(3):
        o = int(off[7]) + s * int(stride[7])
        assert np.array_equal(one[s].view(np.uint32), full[o:o + lens[7]].view(np.uint32))
    with pytest.raises(Exception):
        engine.run(batch)
    batch.free()
    # a second batch of the same shape is served from the pool
    m0 = engine.pool_stats()["cudaMalloc_calls"]
    b2 = engine.upload(a); engine.run(b2); b2.free()
    assert engine.pool_stats()["cudaMalloc_calls"] == m0
