"""Range-sharded dedup: the protocol of openge_b200/sharded.py must give, for any
number of ranks, exactly the flags of the single-stream run.

CPU part: the numpy model of the protocol (tests/sharded_model.py) in-process and under
torch.distributed/gloo with world_size 2 -- this exercises the host orchestration (ranges,
all-to-all exchanges, phase order) without a GPU.  GPU part: the CUDA engine with all ranks as
contexts on one device, against the same oracle flags.
"""
import os
import socket

import numpy as np
import pytest
import torch

import fixtures
import oracle
from conftest import load_golden
from openge_b200 import bamio, sharded, synth
from sharded_model import ModelShardEngine


def straddling_case():
    """Names whose sightings straddle shards in the ways that matter: a supplementary-like third
    record BEFORE its pair (shifts the toggle), after it, cross-contig mates, duplicates of a
    cross-shard pair, and fragments at the same position on both sides of a cut."""
    bam = synth.make("C3", 0.004, seed=11)
    return bam


def straddle_fixture():
    """12 records, cut in two after record 6 (key boundary chr1:5000): every cross-rank mechanism in one file.
      Z     a third sighting (0x800, primary to the reference) on shard 0 BEFORE its couple on shard 1:
            the toggle pairs (Z@1000, Z@5100) and leaves Z@5400 -> shard 1's local couple is retracted
      P, Q  duplicate pairs with one mate on each shard -> published, replayed, marks go to both shards
      A1/A2 reverse-strand duplicates: A1 lies on shard 0 but its unclipped end is in shard 1's key range
      B1/B2 forward duplicates: B2 lies on shard 1 but its unclipped start is in shard 0's key range"""
    text = "@HD\tVN:1.4\tSO:coordinate\n@SQ\tSN:chr1\tLN:100000\n@RG\tID:rg1\tLB:libA\tSM:s\n"
    r = fixtures._rec
    recs = [
        r("Z", 99 | 0x800, 0, 1000, "100M", 0, 5400, "I"),
        r("P", 99, 0, 2000, "100M", 0, 6000, "I"),
        r("Q", 99, 0, 2000, "100M", 0, 6000, "5"),
        r("F", 0, 0, 3000, "100M", -1, 0, "I"),
        r("A1", 16, 0, 4990, "100M", -1, 0, "5"),
        r("B1", 0, 0, 4998, "100M", -1, 0, "I"),
        r("A2", 16, 0, 5000, "10S90M", -1, 0, "I"),
        r("B2", 0, 0, 5003, "5S95M", -1, 0, "5"),
        r("Z", 99, 0, 5100, "100M", 0, 5400, "I"),
        r("Z", 147, 0, 5400, "100M", 0, 5100, "I"),
        r("P", 147, 0, 6000, "100M", 0, 2000, "I"),
        r("Q", 147, 0, 6000, "100M", 0, 2000, "5"),
    ]
    records, offsets = bamio.concat_records(recs)
    return bamio.BamFile(text=text, refs=[("chr1", 100000)], records=records, offsets=offsets)


def test_straddle_fixture_expected_flags():
    bam = straddle_fixture()
    dup = (oracle.markdup(bam.records, bam.offsets, bam.text) & 0x400) != 0
    #        Z    P    Q    F    A1   B1   A2   B2   Z    Z    P    Q
    assert dup.tolist() == [False, False, True, False, True, False, False, True, False, False, False, True]
    if oracle.ref_available():
        try:
            ref = oracle.ref_dedup(bam)
        except oracle.RefHang:
            pytest.skip("reference did not terminate")
        assert np.array_equal((ref.flags() & 0x400) != 0, dup)


def model_flags(bam, world):
    plan, shards = sharded.split_bam(bam, world)
    engines = [ModelShardEngine(rec, off, bam.text, plan, r) for r, (rec, off) in enumerate(shards)]
    info = sharded.run_phases(engines, sharded.LocalExchange())
    return np.concatenate([e.flags() for e in engines]), info


def case_bam(case):
    if case == "c3":
        return straddling_case()
    if case == "straddle":
        return straddle_fixture()
    return load_golden(case)[0]


@pytest.mark.parametrize("world", [1, 2, 3, 5])
@pytest.mark.parametrize("case", ["straddle", "a3_fixture1", "a3_fixture2", "edge_cases", "c3"])
def test_protocol_model_equals_oracle(case, world):
    bam = case_bam(case)
    want = oracle.markdup(bam.records, bam.offsets, bam.text)
    got, info = model_flags(bam, world)
    assert np.array_equal(got, want)
    if world == 2 and case == "straddle":      # every exchange carried something
        assert info["published"] >= 7 and info["routed"] >= 2 and info["marks"] >= 2


def test_split_plan_covers_the_file():
    bam = straddling_case()
    for world in (1, 2, 7):
        plan, shards = sharded.split_bam(bam, world)
        assert plan.bases[0] == 0 and plan.bases[-1] == bam.n and len(shards) == world
        assert sum(len(o) - 1 for _, o in shards) == bam.n
        assert b"".join(r.tobytes() for r, _ in shards) == bam.records.tobytes()
        keys = list(zip(plan.split_ref, plan.split_pos))
        assert keys == sorted(keys, key=lambda k: (k[0] if k[0] >= 0 else 1 << 40, k[1]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out_dir, kind="alltoall"):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bam = straddling_case()
        plan, shards = sharded.split_bam(bam, world)
        rec, off = shards[rank]
        ex = (sharded.GatherExchange if kind == "gather" else sharded.AllToAllExchange)(dist, torch.device("cpu"))
        for it in range(3):      # first run: sized protocol; later runs: one framed collective per exchange
            eng = ModelShardEngine(rec, off, bam.text, plan, rank)
            if it == 2:
                ex.cap = {k: 4096 for k in ex.cap}      # frames too small: every rank must fall back together
            info = sharded.run_phases([eng], ex)
            np.save(os.path.join(out_dir, "flags%d_%d.npy" % (rank, it)), eng.flags())
        np.save(os.path.join(out_dir, "flags%d.npy" % rank), eng.flags())
        np.save(os.path.join(out_dir, "info%d.npy" % rank), np.array([info["published"], info["routed"], info["marks"], ex.bytes_moved,
                                                                      ex.calls["framed"], ex.calls["sized"]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["alltoall", "gather"])
def test_orchestration_under_gloo_world2(tmp_path, kind):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path), kind), nprocs=world, join=True)
    bam = straddling_case()
    want = oracle.markdup(bam.records, bam.offsets, bam.text)
    got = np.concatenate([np.load(tmp_path / ("flags%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got, want)
    for it in range(3):
        got = np.concatenate([np.load(tmp_path / ("flags%d_%d.npy" % (r, it))) for r in range(world)])
        assert np.array_equal(got, want)
    i0, i1 = np.load(tmp_path / "info0.npy"), np.load(tmp_path / "info1.npy")
    assert np.array_equal(i0[:3], i1[:3]) and i0[0] > 0      # both ranks saw the same exchanged lists
    assert i0[4] >= 3 and i0[5] >= 4 and np.array_equal(i0[4:], i1[4:])      # framed calls happened; so did the joint fall-backs


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("case", ["straddle", "a3_fixture1", "a3_fixture2", "edge_cases", "c3", "synth_C1", "synth_C4", "synth_C5"])
def test_cuda_shards_in_process_equal_oracle(case, world):
    bam = case_bam(case)
    want = oracle.markdup(bam.records, bam.offsets, bam.text)
    got, info = sharded.dedup_in_process(bam, world)
    assert np.array_equal(got, want)
    if world == 2 and case == "straddle":
        assert info["published"] >= 7 and info["routed"] >= 2 and info["marks"] >= 2


@pytest.mark.gpu
def test_cuda_route_overflow_sweep(monkeypatch):
    """A routing buffer that is too small: the entries that found no room are collected by a second sweep."""
    monkeypatch.setenv("OGE_ROUTE_CAP", "1")      # read by the -DOGE_TESTING build only
    bam = straddle_fixture()
    want = oracle.markdup(bam.records, bam.offsets, bam.text)
    from openge_b200 import dedup
    with dedup.testing_library():
        got, info = sharded.dedup_in_process(bam, 2)
    assert np.array_equal(got, want) and info["routed"] >= 2


@pytest.mark.gpu
def test_cuda_shards_larger_synthetic():
    bam = synth.make("C3", 0.05, seed=21)
    want = oracle.markdup(bam.records, bam.offsets, bam.text)
    for world in (2, 4):
        got, _ = sharded.dedup_in_process(bam, world)
        assert np.array_equal(got, want)


@pytest.mark.gpu
def test_bench_shards_concatenate_to_one_sorted_file_and_match_oracle():
    """The weak-scaling bench's per-rank generator: the shards form one coordinate-sorted file; the
    sharded run over them equals the oracle on the concatenation."""
    world = 2
    parts = [sharded.make_rank_shard("C2", 0.002, r, world, pinned=False) for r in range(world)]
    text, contigs = parts[0][2], parts[0][3]
    rec = np.concatenate([p[0] for p in parts])
    off = bamio.frame_records(rec.tobytes())
    whole = bamio.BamFile(text=text, refs=contigs, records=rec, offsets=off)
    want = oracle.markdup(whole.records, whole.offsets, whole.text)
    bases = np.cumsum([0] + [len(p[1]) - 1 for p in parts])
    plan = sharded.ShardPlan(bases, [sharded._first_key(parts[1][0], parts[1][1], 0)[0]], [sharded._first_key(parts[1][0], parts[1][1], 0)[1]])
    engines = [sharded.CudaShardEngine(p[0], p[1], text, contigs, plan, r) for r, p in enumerate(parts)]
    info = sharded.run_phases(engines, sharded.LocalExchange())
    got = np.concatenate([e.flags() for e in engines])
    for e in engines:
        e.close()
    assert np.array_equal(got, want)
    assert info["published"] > 0
