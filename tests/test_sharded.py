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


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bam = straddling_case()
        plan, shards = sharded.split_bam(bam, world)
        rec, off = shards[rank]
        ex = sharded.AllToAllExchange(dist, torch.device("cpu"))
        for it in range(2):
            eng = ModelShardEngine(rec, off, bam.text, plan, rank)
            sharded.agree_entry_bytes([eng], dist, torch.device("cpu"))
            info = sharded.run_phases([eng], ex, ranks=[rank], world=world)
            np.save(os.path.join(out_dir, "flags%d_%d.npy" % (rank, it)), eng.flags())
        np.save(os.path.join(out_dir, "flags%d.npy" % rank), eng.flags())
        tot = torch.tensor([info["published"], info["routed"], info["marks"]], dtype=torch.int64)
        dist.all_reduce(tot)
        np.save(os.path.join(out_dir, "info%d.npy" % rank), np.array(tot.tolist() + [ex.bytes_moved, ex.calls]))
    finally:
        dist.destroy_process_group()


def test_orchestration_under_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    bam = straddling_case()
    want = oracle.markdup(bam.records, bam.offsets, bam.text)
    got = np.concatenate([np.load(tmp_path / ("flags%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got, want)
    for it in range(2):
        got = np.concatenate([np.load(tmp_path / ("flags%d_%d.npy" % (r, it))) for r in range(world)])
        assert np.array_equal(got, want)
    i0, i1 = np.load(tmp_path / "info0.npy"), np.load(tmp_path / "info1.npy")
    assert np.array_equal(i0[:3], i1[:3]) and i0[0] > 0      # published, routed, marks over all ranks
    assert i0[4] == 8 and i1[4] == 8                           # four all-to-all exchanges per run, two runs


def test_bench_shards_cut_inside_contigs_and_form_one_sorted_file():
    """The weak-scaling bench's per-rank generator (C5 shape): the cuts fall inside contigs, pairs straddle them, and the
    concatenation of the shards is one coordinate-sorted file the protocol model dedups exactly like the oracle."""
    world = 3
    parts = [sharded.make_rank_shard("C5", 0.00002, r, world, pinned=False) for r in range(world)]
    text, contigs = parts[0][2], parts[0][3]
    rec = np.concatenate([p[0] for p in parts])
    off = bamio.frame_records(rec.tobytes())
    o = off[:-1].astype(np.int64)
    rp = rec[o[:, None] + np.arange(4, 12)].copy().view("<i4").reshape(-1, 2)
    key = rp[:, 0].astype(np.int64) * (1 << 32) + rp[:, 1]
    assert (np.diff(key) >= 0).all()                                  # one coordinate-sorted file
    firsts = [sharded._first_key(p[0], p[1], 0) for p in parts[1:]]
    assert any(pos > 1000 for _, pos in firsts)                       # a cut inside a contig
    names = {}
    for r, p in enumerate(parts):
        for i in range(len(p[1]) - 1):
            q = p[0][int(p[1][i]):int(p[1][i + 1])]
            names.setdefault(q[36:36 + int(q[12]) - 1].tobytes(), set()).add(r)
    assert sum(1 for v in names.values() if len(v) > 1) > 50          # mates on two ranks: straddlers and cross-contig pairs
    whole = bamio.BamFile(text=text, refs=contigs, records=rec, offsets=off)
    want = oracle.markdup(whole.records, whole.offsets, whole.text)
    bases = np.cumsum([0] + [len(p[1]) - 1 for p in parts])
    plan = sharded.ShardPlan(bases, [f[0] for f in firsts], [f[1] for f in firsts])
    engines = [ModelShardEngine(p[0], p[1], text, plan, r) for r, p in enumerate(parts)]
    info = sharded.run_phases(engines, sharded.LocalExchange())
    assert np.array_equal(np.concatenate([e.flags() for e in engines]), want)
    assert info["published"] > 100 and info["marks"] > 0


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
@pytest.mark.parametrize("world", [2, 3])
def test_bench_shards_on_the_cuda_engine_match_oracle(world):
    """The bench's C5-shaped shards (cuts inside contigs, straddling pairs, cross-contig pairs, names with four sightings across
    a cut) through the CUDA engine, all ranks as contexts on one GPU, against the oracle on the concatenation."""
    parts = [sharded.make_rank_shard("C5", 0.0005, r, world, pinned=False) for r in range(world)]
    text, contigs = parts[0][2], parts[0][3]
    rec = np.concatenate([p[0] for p in parts])
    off = bamio.frame_records(rec.tobytes())
    want = oracle.markdup(rec, off, text)
    bases = np.cumsum([0] + [len(p[1]) - 1 for p in parts])
    firsts = [sharded._first_key(p[0], p[1], 0) for p in parts[1:]]
    plan = sharded.ShardPlan(bases, [f[0] for f in firsts], [f[1] for f in firsts])
    engines = [sharded.CudaShardEngine(p[0], p[1], text, contigs, plan, r) for r, p in enumerate(parts)]
    try:
        sharded.agree_entry_bytes(engines)
        info = sharded.run_phases(engines, sharded.LocalExchange())
        got = np.concatenate([e.flags() for e in engines])
    finally:
        for e in engines:
            e.close()
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, "%d flag words differ, first at record %d" % (len(bad), int(bad[0]))
    assert info["published"] > 0 and info["routed"] > 0 and info["marks"] > 0


def _nccl_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    ndev = torch.cuda.device_count()
    local = rank % ndev
    torch.cuda.set_device(local)
    # one GPU per rank where the box has them (NCCL); ranks sharing one GPU cannot use NCCL: gloo moves the (CPU-staged) lists
    backend = "nccl" if ndev >= world else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", local)
        if backend == "nccl":
            res = sharded.parity_pass("C5", rank, world, local, dist, dev, oracle.markdup, reads=400_000)
        else:
            res = _parity_pass_staged(rank, world, local, dist)
        if rank == 0:
            np.save(os.path.join(out_dir, "verdict.npy"), np.array([res["records"], res["mismatches"], res["published"], res["routed"], res["marks"],
                                                                    1 if backend == "nccl" else 0]))
    finally:
        dist.destroy_process_group()


class _StagedExchange(sharded.AllToAllExchange):
    """all_to_all_single over gloo for CUDA engines on a box with fewer GPUs than ranks: the lists are staged through host memory."""

    def __call__(self, outs):
        (mine,) = outs
        cpu = tuple(sharded.Bucketed(sharded.as_tensor(b.data).cpu(), b.counts, b.item) for b in mine)
        res = super().__call__([cpu])
        return tuple([t[0].cuda()] for t in res)


def _parity_pass_staged(rank, world, local, dist):
    cpu = torch.device("cpu")
    scale = 400_000 / (2.0 * 400_000_000) * 8.0 / world
    rec, offs, text, contigs, _ = sharded.make_rank_shard("C5", scale, rank, world, pinned=False)
    n = len(offs) - 1
    mine = torch.tensor([n] + list(sharded._first_key(rec, offs, 0)), dtype=torch.int64)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    counts = [int(v[0]) for v in allv]
    bases = np.cumsum([0] + counts)
    plan = sharded.ShardPlan(bases, [int(v[1]) for v in allv[1:]], [int(v[2]) for v in allv[1:]])
    eng = sharded.CudaShardEngine(rec, offs, text, contigs, plan, rank, device=local)
    try:
        sharded.agree_entry_bytes([eng], dist, cpu)
        info = sharded.run_phases([eng], _StagedExchange(dist, cpu), ranks=[rank], world=world)
        flags = torch.from_numpy(eng.flags().astype(np.int32))
    finally:
        eng.close()
    pad = torch.zeros(max(counts), dtype=torch.int32)
    pad[:n] = flags
    got = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(got, pad)
    tot = torch.tensor([info["published"], info["routed"], info["marks"]], dtype=torch.int64)
    dist.all_reduce(tot)
    if rank != 0:
        return None
    parts = [sharded.make_rank_shard("C5", scale, r, world, pinned=False) for r in range(world)]
    whole = np.concatenate([p[0] for p in parts])
    want = oracle.markdup(whole, bamio.frame_records(whole.tobytes()), text)
    have = np.concatenate([g[:c].numpy().astype(np.uint16) for g, c in zip(got, counts)])
    return {"records": len(want), "mismatches": int((have != want).sum()), "published": int(tot[0]), "routed": int(tot[1]), "marks": int(tot[2])}


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_cuda_engine_one_process_per_rank_over_the_real_exchange(tmp_path, world):
    """One process per rank, the CUDA engine and torch.distributed's all_to_all_single between them -- NCCL when the box has a
    GPU per rank, gloo (lists staged through the host) when the ranks have to share one -- on a C5-shaped file with cuts inside
    contigs and keys longer than a name tag's 29 bytes; the flags gathered on rank 0 must equal the oracle's on the whole file."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_nccl_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    v = np.load(tmp_path / "verdict.npy")
    assert v[0] > 300_000 and v[1] == 0, "%d of %d flag words differ from the oracle" % (v[1], v[0])
    assert v[2] > 0 and v[3] > 0
