"""Range-sharded dedup over several GPUs: host orchestration.

One coordinate-sorted file is cut into contiguous record ranges, one per rank; the cuts may fall anywhere, inside
contigs included.  Each rank keeps its records, end entries, sorts and selects on its own GPU.  What crosses ranks are
small lists, and each goes only where it is needed -- an all-to-all with uneven splits per exchange (NCCL
all_to_all_single on GPUs, gloo in the CPU tests), four per run:

    begin   ->  published entries (records whose RG:name key was not seen exactly twice on the rank)  -> the rank that
                OWNS the name (key hash mod world), which alone replays it; their 8-byte key hashes -> every other rank;
                copies of the fragment ends whose key lies in another rank's coordinate range -> that rank
    probe   ->  a local pair of a name published elsewhere is retracted, its two records published (round 2) -> the
                name's owner; local pair ends whose key lies in another rank's range -> that rank; the fragment
                sort/select start on a side stream
    replay  ->  the owner replays its names over all ranks' sightings in global file order; pairs whose key range it
                does not own -> the rank that does
    finish  ->  pair sort/select over what the rank owns; marks -> the rank that holds the record
    apply       flag write

The result equals the reference's single-stream `openge dedup --nosplit -v`, not its own split-by-chromosome mode
(SURVEY F2).  The per-phase device work is behind `ShardEngine`; `CudaShardEngine` drives the C ABI
(oge_gpu_shard_*), the test suite plugs in a numpy model of the same protocol (tests/sharded_model.py) to run the
orchestration under gloo without a GPU.
"""
from __future__ import annotations

import time

import numpy as np

ROUTE_BYTES, MARK_BYTES, HASH_BYTES = 32, 4, 8


# --------------------------------------------------------------------------------------- ranges
class ShardPlan:
    """Record ranges and key ranges of a `world`-way split."""

    def __init__(self, bases, split_ref, split_pos):
        self.bases = [int(b) for b in bases]                  # world + 1 global ordinals
        self.split_ref = [int(r) for r in split_ref]          # world - 1: refID of the first record of ranks 1..
        self.split_pos = [int(p) for p in split_pos]
        self.world = len(self.bases) - 1
        self.global_n = self.bases[-1]


def _first_key(records, offsets, i):
    o = int(offsets[i])
    ref, pos = np.frombuffer(records[o + 4: o + 12].tobytes(), dtype="<i4")
    return int(ref), int(pos)


def split_bam(bam, world: int):
    """Cut a BamFile into `world` contiguous record ranges of (nearly) equal record count.
    -> (ShardPlan, [(records, offsets), ...])"""
    n = bam.n
    cuts = [n * r // world for r in range(world + 1)]
    shards, sref, spos = [], [], []
    for r in range(world):
        lo, hi = cuts[r], cuts[r + 1]
        b0, b1 = int(bam.offsets[lo]), int(bam.offsets[hi])
        shards.append((bam.records[b0:b1], (bam.offsets[lo: hi + 1] - bam.offsets[lo]).astype(np.uint64)))
        if r > 0:
            if lo < n:
                ref, pos = _first_key(bam.records, bam.offsets, lo)
            else:
                ref, pos = -1, -1      # an empty tail shard owns no key
            sref.append(ref)
            spos.append(pos)
    return ShardPlan(cuts, sref, spos), shards


# --------------------------------------------------------------------------------------- lists that cross ranks
class Bucketed:
    """What a phase hands to an exchange: items of `item` bytes ordered by destination rank, counts[d] of them for rank d.
    `data` is a 1-D uint8 tensor or a DevChunk."""

    def __init__(self, data, counts, item):
        self.data, self.counts, self.item = data, [int(c) for c in counts], int(item)

    def segment(self, d):
        t = as_tensor(self.data)
        lo = sum(self.counts[:d]) * self.item
        return t[lo: lo + self.counts[d] * self.item]


class ShardEngine:
    """Per-rank device work between the exchanges.  Inputs are 1-D uint8 torch tensors on `device`."""
    device = "cpu"
    entry_bytes = 64

    def key_bytes(self): raise NotImplementedError                    # longest RG:name key of the rank's records
    def set_entry_bytes(self, nbytes): raise NotImplementedError
    def begin(self): raise NotImplementedError                        # -> (pub, hashes tensor, frag_route)
    def probe(self, hashes_in, frag_route_in): raise NotImplementedError   # -> (pub2, pair_route)
    def replay(self, pub_in): raise NotImplementedError               # -> owner_route
    def finish(self, pair_route_in): raise NotImplementedError        # -> marks
    def apply(self, marks_in): raise NotImplementedError
    def flags(self): raise NotImplementedError


class _DevView:
    """Borrowed device memory as a __cuda_array_interface__ object (torch.as_tensor makes a view)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class DevChunk:
    """A list a shard call handed out: device pointer + size, valid until the next call on that context."""

    def __init__(self, engine, ptr, nbytes):
        self.engine, self.ptr, self.nbytes = engine, ptr, int(nbytes)
        self._t = None

    def numel(self):
        return self.nbytes

    def tensor(self):
        if self._t is None:
            torch = self.engine.torch
            if not self.nbytes:
                self._t = torch.empty(0, dtype=torch.uint8, device=self.engine.device)
            else:
                with torch.cuda.device(self.engine.device):
                    self._t = torch.as_tensor(_DevView(self.ptr, self.nbytes), device=self.engine.device).clone()
        return self._t


def as_tensor(x):
    return x.tensor() if isinstance(x, DevChunk) else x


class CudaShardEngine(ShardEngine):
    """One oge_gpu_dedup_ctx driven through the sharded C ABI (include/oge_gpu_dedup.h)."""

    def __init__(self, records, offsets, header_text, refs, plan: ShardPlan, rank: int, device: int = 0, pinned_ptr=None,
                 profile_events=False):
        import torch
        from . import dedup
        self.torch = torch
        self.device = "cuda:%d" % device
        self.rank, self.world = rank, plan.world
        self.n = len(offsets) - 1
        max_len = max([l for _, l in refs], default=0)
        self.ctx = dedup.DedupContext(n_ref=len(refs), max_ref_len=max_len, device=device, rank=rank, world=plan.world,
                                      index_base=plan.bases[rank], profile_events=profile_events,
                                      capacity_records=self.n, capacity_bytes=int(len(records)))
        self.ctx.set_header(header_text)
        self.ctx.shard_setup(plan.global_n, plan.bases, plan.split_ref, plan.split_pos)
        if self.n:
            if pinned_ptr is not None:
                self._off_pin = dedup.PinnedBuffer(offsets.nbytes)
                self._off_pin.array.view(np.uint64)[:] = offsets
                self.ctx.push_async(pinned_ptr, int(len(records)), self._off_pin.ptr, self.n)
                self.ctx.sync()
            else:
                self.ctx.push(records, offsets)

    def key_bytes(self):
        return self.ctx.shard_key_bytes()

    def set_entry_bytes(self, nbytes):
        self.entry_bytes = int(nbytes)
        self.ctx.shard_set_entry_bytes(nbytes)

    def _bucketed(self, ptr, counts, item):
        return Bucketed(DevChunk(self, ptr, sum(counts) * item), counts, item)

    def _give(self, t, item):
        torch = self.torch
        assert t.dtype == torch.uint8 and t.numel() % item == 0
        torch.cuda.current_stream(self.device).synchronize()      # the library works on its own stream
        return (t.data_ptr() if t.numel() else None), t.numel() // item

    def init_nccl(self, dist, device):
        """Join the library's own NCCL communicator: rank 0 makes the id, torch.distributed carries its 128 bytes."""
        import torch
        t = torch.zeros(128, dtype=torch.uint8, device=device)
        if dist.get_rank() == 0:
            t = torch.frombuffer(bytearray(self.ctx.shard_comm_id()), dtype=torch.uint8).to(device)
        dist.broadcast(t, src=0)
        self.ctx.shard_comm_init(bytes(t.cpu().numpy().tobytes()))

    def step(self):
        """One whole run inside the library: phases and NCCL exchanges, no Python in between."""
        return self.ctx.shard_step()

    def begin(self):
        (pp, pc), (hp, nh), (fp, fc) = self.ctx.shard_begin()
        return self._bucketed(pp, pc, self.entry_bytes), DevChunk(self, hp, nh * HASH_BYTES).tensor(), self._bucketed(fp, fc, ROUTE_BYTES)

    def probe(self, hashes_in, frag_route_in):
        (pp, pc), (rp, rc) = self.ctx.shard_probe(*self._give(hashes_in, HASH_BYTES), *self._give(frag_route_in, ROUTE_BYTES))
        return self._bucketed(pp, pc, self.entry_bytes), self._bucketed(rp, rc, ROUTE_BYTES)

    def replay(self, pub_in):
        p, cnt = self.ctx.shard_replay(*self._give(pub_in, self.entry_bytes))
        return self._bucketed(p, cnt, ROUTE_BYTES)

    def finish(self, pair_route_in):
        p, cnt = self.ctx.shard_finish(*self._give(pair_route_in, ROUTE_BYTES))
        return self._bucketed(p, cnt, MARK_BYTES)

    def apply(self, marks_in):
        self.ctx.shard_apply(*self._give(marks_in, MARK_BYTES))

    def flags(self):
        return self.ctx.flags()

    def stats(self):
        return self.ctx.stats()

    def close(self):
        self.ctx.close()


def agree_entry_bytes(engines, dist=None, device=None):
    """Every rank's published entries carry the whole key: 32 bytes + the longest key of ANY rank, rounded up to 32."""
    k = max([e.key_bytes() for e in engines], default=0)
    if dist is not None:
        import torch
        t = torch.tensor([k], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k = int(t.item())
    nbytes = 32 + max(32, (k + 31) // 32 * 32)
    for e in engines:
        e.set_entry_bytes(nbytes)
    return nbytes


# --------------------------------------------------------------------------------------- exchanges
class LocalExchange:
    """All ranks live in this process (several contexts on one GPU, or the numpy model): rank d receives, per list,
    the segments every rank addressed to it, in rank order.  outs = one tuple of Bucketed per rank."""

    def __init__(self):
        self.bytes_moved = 0

    def new_step(self):
        pass

    def __call__(self, outs):
        import torch
        W, k = len(outs), len(outs[0])
        res = []
        for j in range(k):
            per_dest = []
            for d in range(W):
                parts = [outs[r][j].segment(d) for r in range(W)]
                self.bytes_moved += sum(int(p.numel()) for r, p in enumerate(parts) if r != d)
                per_dest.append(torch.cat(parts) if W > 1 else parts[0])
            res.append(per_dest)
        return tuple(res)


class AllToAllExchange:
    """One rank per process under torch.distributed: every exchange is an all-to-all with uneven splits (the path's
    only collective) -- the item counts first, then ONE payload that carries all lists of the exchange, per destination
    [list 0 | list 1 | ...].  outs = [tuple of Bucketed] (this rank's); returns, per list, [received tensor]."""

    def __init__(self, dist, device, timed=False):
        import torch
        self.dist, self.torch, self.device = dist, torch, device
        self.world = dist.get_world_size()
        self.bytes_moved = 0
        self.ms = 0.0
        self.timed = timed and str(device).startswith("cuda")
        self.calls = 0

    def new_step(self):
        pass

    def __call__(self, outs):
        torch, dist, W = self.torch, self.dist, self.world
        (mine,) = outs
        k = len(mine)
        if self.timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        items = [b.item for b in mine]
        send_counts = torch.tensor([[b.counts[d] for b in mine] for d in range(W)], dtype=torch.int64, device=self.device)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts)
        rc = recv_counts.tolist()                                    # rc[r][j] = items of list j rank r addressed to this rank
        in_split = [sum(mine[j].counts[d] * items[j] for j in range(k)) for d in range(W)]
        out_split = [sum(rc[r][j] * items[j] for j in range(k)) for r in range(W)]
        parts = [mine[j].segment(d) for d in range(W) for j in range(k)]
        send = torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device=self.device)
        recv = torch.empty(sum(out_split), dtype=torch.uint8, device=self.device)
        dist.all_to_all_single(recv, send, output_split_sizes=out_split, input_split_sizes=in_split)
        res, base = [], []
        pos = 0
        for r in range(W):
            base.append(pos)
            pos += out_split[r]
        for j in range(k):
            pieces = []
            for r in range(W):
                o = base[r] + sum(rc[r][i] * items[i] for i in range(j))
                pieces.append(recv[o: o + rc[r][j] * items[j]])
            res.append([torch.cat(pieces) if W > 1 else pieces[0]])
        if self.timed:
            e1.record()
            e1.synchronize()
            self.ms += e0.elapsed_time(e1)
        self.bytes_moved += int(send.numel()) - in_split[dist.get_rank()]
        self.calls += 1
        return tuple(res)


# --------------------------------------------------------------------------------------- the protocol
def _hash_bucket(hashes, rank, world):
    """The key hashes of a rank's published entries go to every OTHER rank: the same bytes once per destination."""
    n = int(hashes.numel()) // HASH_BYTES
    counts = [0 if d == rank else n for d in range(world)]
    data = hashes.repeat(world - 1) if world > 1 and n else hashes[:0]
    return Bucketed(data, counts, HASH_BYTES)


def run_phases(engines, exchange, ranks=None, world=None):
    """Drive one sharded run over the local `engines` (one per process under torch.distributed, or all ranks
    in-process): four exchanges.  Afterwards every engine's flags() are final."""
    import torch
    exchange.new_step()
    W = world or len(engines)
    ranks = ranks if ranks is not None else list(range(len(engines)))
    cat = lambda a, b: torch.cat([a, b]) if b.numel() else a      # noqa: E731
    b = [e.begin() for e in engines]
    pub_in, fr_in, hash_in = exchange([(o[0], o[2], _hash_bucket(o[1], r, W)) for o, r in zip(b, ranks)])
    p = [e.probe(h, f) for e, h, f in zip(engines, hash_in, fr_in)]
    pub2_in, pr_in = exchange(p)
    o = [e.replay(cat(a, a2)) for e, a, a2 in zip(engines, pub_in, pub2_in)]
    (or_in,) = exchange([(x,) for x in o])
    m = [e.finish(cat(a, a2)) for e, a, a2 in zip(engines, pr_in, or_in)]
    (m_in,) = exchange([(x,) for x in m])
    for e, t in zip(engines, m_in):
        e.apply(t)
    E = engines[0].entry_bytes if engines else 64
    return {"published": sum(int(a.numel()) + int(a2.numel()) for a, a2 in zip(pub_in, pub2_in)) // E,
            "routed": sum(int(a.numel()) + int(c.numel()) + int(d.numel()) for a, c, d in zip(fr_in, pr_in, or_in)) // ROUTE_BYTES,
            "marks": sum(int(t.numel()) for t in m_in) // MARK_BYTES, "entry_bytes": E}


def dedup_in_process(bam, world: int, device: int = 0):
    """All `world` ranks as contexts on ONE GPU (tests): -> (flags of the whole file, info dict)."""
    plan, shards = split_bam(bam, world)
    engines = [CudaShardEngine(rec, off, bam.text, bam.refs, plan, r, device=device) for r, (rec, off) in enumerate(shards)]
    try:
        agree_entry_bytes(engines)
        info = run_phases(engines, LocalExchange())
        flags = np.concatenate([e.flags() for e in engines]) if engines else np.zeros(0, np.uint16)
        info["stats"] = [e.stats() for e in engines]
    finally:
        for e in engines:
            e.close()
    return flags, info


# --------------------------------------------------------------------------------------- bench support
def _genome_slices(contigs, world):
    """Equal slices of the concatenated genome, one per rank: -> per rank a list of pieces (contig index, start, length).
    The cuts fall wherever they fall -- inside contigs, as the cuts of a real range-sharded file do."""
    total = sum(l for _, l in contigs)
    cuts = [total * r // world for r in range(world + 1)]
    out = []
    for r in range(world):
        lo, hi, pos, pieces = cuts[r], cuts[r + 1], 0, []
        for ci, (_, ln) in enumerate(contigs):
            a, b = max(lo, pos), min(hi, pos + ln)
            if b - a >= 4000:      # a sliver is not worth a piece (the generator needs room for an insert)
                pieces.append((ci, a - pos, b - a))
            pos += ln
        out.append(pieces)
    return out


def _straddlers(contigs, slices, read_len, seed):
    """Pairs whose two reads lie on either side of a cut between two ranks (what a 300-bp insert does at every cut of a real
    file), plus three-sighting names across a cut: a few hundred records per cut, built identically on every rank."""
    from . import bamio
    rng = np.random.default_rng(seed)
    recs = []
    L = read_len
    seq = ("ACGT" * (L // 4 + 1))[:L]
    for r in range(1, len(slices)):
        if not slices[r] or not slices[r - 1]:
            continue
        ci, start, _ = slices[r][0]
        if start == 0:
            continue      # the cut coincides with a contig start: nothing can straddle it
        for k in range(160):
            ins = int(rng.integers(L + 20, 520))
            p1 = start - int(rng.integers(1, ins - 10))        # first read starts before the cut ...
            p2 = p1 + ins - L                                  # ... its mate (reverse strand) ends behind it
            if p1 < 0 or p2 < start:
                p2 = max(p2, start)
            name = "cut%d_%05d" % (r, k // (2 if k % 8 == 0 else 1))      # every eighth pair is a copy of its predecessor's name: 4 sightings
            q1, q2 = int(rng.integers(20, 41)), int(rng.integers(20, 41))
            tags = bamio.tag_z("RG", "rg1")
            recs.append((ci, p1, bamio.build_record(name, 99, ci, p1, 60, "%dM" % L, ci, p2, ins, seq, q1, tags)))
            recs.append((ci, p2, bamio.build_record(name, 147, ci, p2, 60, "%dM" % L, ci, p1, -ins, seq, q2, tags)))
            if k % 5 == 0:      # duplicates of the straddling pair: same ends, other name
                recs.append((ci, p1, bamio.build_record(name + "d", 99, ci, p1, 60, "%dM" % L, ci, p2, ins, seq, q2, tags)))
                recs.append((ci, p2, bamio.build_record(name + "d", 147, ci, p2, 60, "%dM" % L, ci, p1, -ins, seq, q1, tags)))
    recs.sort(key=lambda t: (t[0], t[1]))
    return bamio.concat_records([t[2] for t in recs])


def make_rank_shard(workload, scale, rank, world, pinned=True):
    """Synthetic shard of rank `rank` of a range-sharded file (weak scaling: every rank holds 1/8 of the workload at the given
    scale, so that 8 ranks hold the whole of it).  The genome is the workload's own contig set; rank r draws its reads on
    slice r of the concatenated genome -- the cuts fall inside contigs -- with its own seed; every rank adds the records of
    two small overlays that fall into its slice, generated identically everywhere: pairs whose mates lie on different
    contigs (0.5 % of the pairs: most of them cross ranks) and pairs that straddle the cuts.  The concatenation of the
    shards is one coordinate-sorted file.  -> (records, offsets, header text, contigs, keepalive)"""
    from . import dedup, synth
    cfg, contigs, rgs = synth.config(workload, scale)
    per_rank = max(64, int(cfg.n_templates) // 8)
    slices = _genome_slices(contigs, world)
    hold = {}

    def alloc(tag):
        def f(nbytes):
            if pinned:
                hold[tag] = dedup.PinnedBuffer(nbytes + 64)
                return hold[tag].array[:nbytes]
            return np.empty(nbytes + 64, dtype=np.uint8)[:nbytes]
        return f

    pieces = slices[rank]
    main = synth.SynthCfg.from_buffer_copy(bytes(cfg))
    main.seed = cfg.seed * 1000 + rank
    main.n_templates = per_rank
    main.n_contigs = len(pieces)
    for i, (_, _, ln) in enumerate(pieces):
        main.contig_len[i] = ln
    main.contig_lo = main.contig_hi = 0
    main.cross_contig_frac = 0.0
    rec, offs = synth.generate(main, records_out=alloc("main") if world == 1 else None)
    synth.remap_pieces(rec, offs, [p[0] for p in pieces], [p[1] for p in pieces])
    if world > 1:
        lo = (pieces[0][0], pieces[0][1])
        hi = (pieces[-1][0], pieces[-1][1] + pieces[-1][2])
        over = synth.SynthCfg.from_buffer_copy(bytes(cfg))
        over.seed = cfg.seed * 7919 + 17
        over.n_templates = max(16, int(per_rank * world * 0.005))
        over.cross_contig_frac, over.dup_frac = 1.0, 0.10
        over.single_frac = over.mate_unmapped_frac = over.unmapped_pair_frac = over.secondary_frac = over.supplementary_frac = 0.0
        orec, ooffs = synth.generate(over)
        srec, soffs = _straddlers(contigs, slices, int(cfg.read_len), cfg.seed * 31 + 7)
        # the two small overlays are merged with each other first, so that the shard's 30 GB are copied once, not twice
        none = np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64)
        orec, ooffs = synth.merge_sorted(none[0], none[1], orec, ooffs, synth.records_in_range(orec, ooffs, lo, hi))
        orec, ooffs = synth.merge_sorted(orec, ooffs, srec, soffs, synth.records_in_range(srec, soffs, lo, hi))
        rec, offs = synth.merge_sorted(rec, offs, orec, ooffs, np.ones(len(ooffs) - 1, dtype=np.uint8), records_out=alloc("merged"))
    return rec, offs, synth.header_text(contigs, rgs), contigs, hold


def parity_pass(workload, rank, world, local_rank, dist, dev, checker, reads=2_000_000, native=True):
    """Untimed: a `reads`-record file of the same shape, sharded over the same ranks through the same NCCL exchanges, its
    flags gathered on rank 0 and compared record by record with `checker(records, offsets, header text) -> flags` run over
    the whole file (the caller -- bench.py, the tests -- supplies the CPU oracle; rank 0 regenerates every rank's shard: the
    generator is deterministic).  -> verdict dict on rank 0, None elsewhere."""
    import torch
    from . import synth
    cfg, _, _ = synth.config(workload, 1.0)
    scale = max(1e-6, reads / (2.0 * int(cfg.n_templates)) * 8.0 / world)
    rec, offs, text, contigs, hold = make_rank_shard(workload, scale, rank, world, pinned=False)
    n = len(offs) - 1
    mine = torch.tensor([n] + list(_first_key(rec, offs, 0) if n else (-1, -1)), dtype=torch.int64, device=dev)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    counts = [int(v[0]) for v in allv]
    bases = [0]
    for c in counts:
        bases.append(bases[-1] + c)
    plan = ShardPlan(bases, [int(v[1]) for v in allv[1:]], [int(v[2]) for v in allv[1:]])
    eng = CudaShardEngine(rec, offs, text, contigs, plan, rank, device=local_rank)
    try:
        agree_entry_bytes([eng], dist, dev)
        if native:
            eng.init_nccl(dist, dev)
            info = eng.step()
        else:
            info = run_phases([eng], AllToAllExchange(dist, dev), ranks=[rank], world=world)
        tot = torch.tensor([info["published"], info["routed"], info["marks"]], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        info = {"published": int(tot[0]), "routed": int(tot[1]), "marks": int(tot[2])}
        flags = torch.from_numpy(eng.flags().astype(np.int32)).to(dev)
    finally:
        eng.close()
    pad = torch.zeros(max(counts), dtype=torch.int32, device=dev)
    pad[:n] = flags
    got = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(got, pad)
    if rank != 0:
        return None
    parts = [make_rank_shard(workload, scale, r, world, pinned=False) for r in range(world)]
    whole = np.concatenate([p[0] for p in parts])
    wo = [np.zeros(1, np.uint64)]
    acc = 0
    for p in parts:
        wo.append(p[1][1:] + np.uint64(acc))
        acc += int(p[1][-1])
    want = checker(whole, np.concatenate(wo), text)
    have = np.concatenate([g[:c].cpu().numpy().astype(np.uint16) for g, c in zip(got, counts)])
    return {"checked": True, "records": int(len(want)), "mismatches": int((have != want).sum()), "published": info["published"],
            "routed": info["routed"], "marks": info["marks"],
            "through": ("the library's own NCCL exchanges (grouped ncclSend/ncclRecv), %d ranks" if native else "torch.distributed all_to_all_single (NCCL), %d ranks") % world}


def bench(args, rank, world, local_rank, rec, offs, text, contigs, metric, workload_name):
    """bench.py's N > 1 arm: one rank per process, NCCL all-to-all exchanges, device-timed phases."""
    import json
    import os
    import sys

    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench as benchmod
    from . import dedup

    dev = torch.device("cuda", local_rank)
    n = len(offs) - 1
    # ranges: record counts and first keys of every rank (setup, not timed)
    mine = torch.tensor([n] + list(_first_key(rec, offs, 0) if n else (-1, -1)), dtype=torch.int64, device=dev)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    counts = [int(v[0]) for v in allv]
    bases = [0]
    for c in counts:
        bases.append(bases[-1] + c)
    plan = ShardPlan(bases, [int(v[1]) for v in allv[1:]], [int(v[2]) for v in allv[1:]])

    # untimed: the same path on a small file of the same shape, record by record against the oracle
    parity = None
    if not args.no_oracle_check:
        parity = parity_pass(args.workload, rank, world, local_rank, dist, dev, benchmod.oracle_flags)

    eng = CudaShardEngine(rec, offs, text, contigs, plan, rank, device=local_rank, pinned_ptr=rec.ctypes.data, profile_events=True)
    entry_bytes = agree_entry_bytes([eng], dist, dev)
    native = os.environ.get("OGE_EXCHANGE", "native") != "torch"      # A/B: the exchanges through torch.distributed instead
    ex = AllToAllExchange(dist, dev, timed=True)
    if native:
        eng.init_nccl(dist, dev)

    def one_run():
        return eng.step() if native else run_phases([eng], ex, ranks=[rank], world=world)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step():
        # every phase ends with a host-side sync on the library's streams, so events on torch's stream
        # around the whole step see all of it: device phases, exchanges, and whatever overlaps them
        ex.ms = 0.0
        e0.record()
        info = one_run()
        e1.record()
        e1.synchronize()
        st = eng.stats()
        return e0.elapsed_time(e1), st, info, ex.ms

    sampler = benchmod.ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()
    t0 = time.perf_counter()
    ms, sts, launches, ex_ms = [], [], 0, 0.0
    for _ in range(args.steps):
        m, st, info, xm = step()
        ms.append(m)
        sts.append(st)
        ex_ms += xm
        launches += st["launches"]
    dist.barrier()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    sampler.mark_end()
    clocks = sampler.stop()

    # max over ranks of the event-timed step
    t = torch.tensor([float(np.mean(ms)), ex_ms / args.steps, float(st["n_duplicates"]), float(n), wall_ms,
                      float(info["published"]), float(info["routed"]), float(info["marks"])], dtype=torch.float64, device=dev)
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    ms_per_step, total_reads, total_dups = float(mx[0]), int(sm[3]), int(sm[2])

    # end to end from pinned host buffers: push + phases + flags back
    flags_pin = dedup.PinnedBuffer(max(2, n * 2))
    e2e = []
    for it in range(1 + max(1, min(args.steps, 2))):
        eng.ctx.reset()
        dist.barrier()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if n:
            eng.ctx.push_async(rec.ctypes.data, int(rec.nbytes), eng._off_pin.ptr, n)
        one_run()
        eng.ctx.flags(flags_pin.array.view(np.uint16)[:n])
        dist.barrier()
        e2e.append(time.perf_counter() - t1)
    e2e_s = torch.tensor([float(np.mean(e2e[1:]))], dtype=torch.float64, device=dev)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = torch.tensor([float(rec.nbytes + offs.nbytes), float(n * 2)], dtype=torch.float64, device=dev)
    dist.all_reduce(h2d, op=dist.ReduceOp.SUM)

    if rank == 0:
        peak, peak_src = benchmod.load_peaks()
        mean = lambda f: float(np.mean([f(s) for s in sts]))      # noqa: E731
        k1_ms = mean(lambda s: s["ms_kernel"]["endbuild"])
        a_parse = benchmod.parse_bytes_per_read(rec, offs, n)
        n_pe = 2 * st["n_local_pairs"] + st["n_join_leftovers"]
        k1_bytes = n * a_parse + st["n_frag_entries"] * 16 + n_pe * 24
        pass_ms, pass_bytes, pass_launches = mean(lambda s: s["ms_sort_pass_kernels"]), mean(lambda s: s["sort_pass_bytes"]), mean(lambda s: s["sort_pass_launches"])
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sb = benchmod.sample_bam(args.workload, 1_000_000)
            secs, kind = benchmod.cpu_reference_run(sb, cores, reps=1, timeout=300)
            cpu = {"value": sb.n / secs[0], "unit": "reads/s", "cores": cores if kind == "reference" else 1, "kind": kind,
                   "sample": "%d reads of workload %s (same generator, scaled), one process on rank 0's host cores" % (sb.n, args.workload)}
        line = {
            "metric": metric, "value": total_reads / (ms_per_step * 1e-3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": "%s, range-sharded over %d ranks: rank r holds slice r of the concatenated genome (the cuts fall inside "
                                   "contigs), 1/8 of the workload's reads per rank; pairs straddle every cut, 0.5%% of the pairs have "
                                   "their mates on different contigs" % (workload_name, world),
                       "reads": total_reads, "reads_rank0": n, "l2": "inputs larger than L2",
                       "wall_ms_per_step": float(mx[4]), "step_ms_rank0": [float(x) for x in ms], "exchange_ms_per_step": float(mx[1]), "device_phase_ms_rank0": st["ms_total"],
                       "timing": "CUDA events around each step (phases sync the library streams before returning), max over ranks",
                       "duplicates_flagged": total_dups,
                       "published_entries": int(sm[5]), "routed_entries": int(sm[6]), "marks_exchanged": int(sm[7]), "published_entry_bytes": entry_bytes,
                       "stage_ms_rank0": {k: st[k] for k in ("ms_endbuild", "ms_join", "ms_sort_pair", "ms_sort_frag", "ms_select", "ms_flags")},
                       "parity_vs_oracle": parity, "host_numa_binding_rank0": getattr(args, "numa", None),
                       "exchanges": "inside the library: grouped ncclSend/ncclRecv per peer on its own stream" if native else "torch.distributed all_to_all_single",
                       "parallelism": "range-sharded x%d; four all-to-all exchanges (NCCL, uneven splits) of small lists per "
                                      "step: published entries to the name's owner (hash mod %d) + their hashes to all + boundary fragment ends, "
                                      "round-2 entries + local pair ends, replayed pair ends, marks" % (world, world)},
            "e2e": {"value": total_reads / float(e2e_s[0]), "unit": "reads/s", "h2d_bytes_per_step": int(h2d[0]),
                    "d2h_bytes_per_step": int(h2d[1]), "ms_per_step": float(e2e_s[0]) * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "endbuild_kernel (K1 end-build on rank 0, the largest stage of the step)",
                         "achieved": (k1_bytes / 1e9) / (k1_ms * 1e-3) if k1_ms > 0 else None, "peak": peak, "unit": "GB/s",
                         "frac": (k1_bytes / 1e9) / (k1_ms * 1e-3) / peak if k1_ms > 0 else None, "peak_source": peak_src, "traffic": None,
                         "avg_launch_ms": k1_ms, "algorithmic_bytes_per_launch": k1_bytes, "rank": 0,
                         "stages": [{"stage": "K3 sort pass", "kernel": "rs_pass_v2", "ms_per_step": pass_ms,
                                     "achieved": (pass_bytes / 1e9) / (pass_ms * 1e-3) if pass_ms > 0 else None,
                                     "frac": (pass_bytes / 1e9) / (pass_ms * 1e-3) / peak if pass_ms > 0 else None,
                                     "launches_per_step": pass_launches}]},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0
