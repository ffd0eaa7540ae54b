"""Multi-GPU host logic on the CPU: the bundle partitioner (stg_shard_assign through the C ABI) and the one
collective of the path (sum of the global totals), world_size 2 over gloo.  Each rank runs the ORACLE on its shard
(there is no GPU here); the reduced totals and the per-bundle results must equal the single-process ones."""
import os
import socket
import sys

import numpy as np
import pytest

from stringtie_b200 import shard, synth
from tests import orc
from tests.conftest import ROOT


def _batch():
    return synth.make_batch(n_bundles=40, reads_per_bundle=300, seed=5)


def test_assign_is_a_balanced_partition():
    a = _batch()
    for world in (1, 2, 3, 8):
        parts, cost = shard.shard(a, world)
        allb = np.sort(np.concatenate(parts))
        assert np.array_equal(allb, np.arange(a["b_start"].size))           # complete and disjoint
        assert all(np.all(np.diff(p) > 0) for p in parts if p.size > 1)      # genomic order kept inside a rank
        reads = np.diff(a["b_read_off"].astype(np.int64))
        span = a["b_end"].astype(np.int64) - a["b_start"] + 3
        c = reads + span / 64.0
        assert np.allclose(cost, [c[p].sum() for p in parts])
        assert cost.max() - cost.min() <= c.max() + 1e-9                     # greedy LPT bound


def test_take_bundles_round_trip():
    a = _batch()
    idx = np.array([3, 4, 10, 39, 0])
    sub = shard.take_bundles(a, idx)
    cov, off, good, left, right = orc.run(sub)
    fcov, foff, fgood, fleft, fright = orc.run(a)
    for k, b in enumerate(idx):
        assert np.array_equal(cov[off[k]:off[k + 1]].view(np.uint32), fcov[foff[b]:foff[b + 1]].view(np.uint32))
        j0, j1 = int(a["b_junc_off"][b]), int(a["b_junc_off"][b + 1])
        s0 = int(sub["b_junc_off"][k])
        assert np.array_equal(good[s0:s0 + j1 - j0], fgood[j0:j1])
        assert np.array_equal(left[s0:s0 + j1 - j0], fleft[j0:j1])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = _batch()
    parts, _ = shard.shard(a, world)
    mine = shard.take_bundles(a, parts[rank])
    cov, off, good, left, right = orc.run(mine)
    lens, _ = orc.cov_layout(mine)
    tot = np.zeros(3)
    for b in range(lens.size):
        tot[0] += orc.get_cov(cov[off[b]:off[b + 1]], int(lens[b]), 1, 0, int(lens[b]) - 2)
    tot[1] = float(mine["r_start"].size)
    tot[2] = float(lens.size)
    # the reference's global sums Num_Fragments / Frag_Len (stringtie.cpp:1490-1497) come out of rows a8 + a9
    gr = orc.groups(mine, cov, off, good, left, right)
    tot = np.concatenate([tot, [sum(g["num_fragments"] for g in gr), sum(g["frag_len"] for g in gr)]])
    red = shard.allreduce_totals(tot, dist)
    q.put((rank, red.tolist(), parts[rank].tolist(), float(good.sum())))
    dist.destroy_process_group()


def test_two_ranks_gloo_allreduce_matches_single_process():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    a = _batch()
    cov, off, good, left, right = orc.run(a)
    lens, _ = orc.cov_layout(a)
    want = sum(orc.get_cov(cov[off[b]:off[b + 1]], int(lens[b]), 1, 0, int(lens[b]) - 2) for b in range(lens.size))
    full = orc.groups(a, cov, off, good, left, right)
    want_nf, want_fl = sum(g["num_fragments"] for g in full), sum(g["frag_len"] for g in full)
    assert want_nf > 0
    for rank, red, part, gsum in res:
        assert abs(red[0] - want) <= 1e-9 * max(1.0, want)
        assert red[1] == a["r_start"].size and red[2] == lens.size
        assert abs(red[3] - want_nf) <= 1e-9 * max(1.0, want_nf) and abs(red[4] - want_fl) <= 1e-9 * max(1.0, want_fl)
    assert abs(sum(r[3] for r in res) - good.sum()) < 1e-9
    assert sorted(res[0][2] + res[1][2]) == list(range(lens.size))
