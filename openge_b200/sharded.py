"""Range-sharded dedup over several GPUs: host orchestration.

One coordinate-sorted file is cut into contiguous record ranges, one per rank.  Each rank keeps its
records, end entries, sorts and selects on its own GPU; small lists cross ranks in three exchanges per
run, each delivering every rank's lists to every rank with an all-to-all (NCCL on GPUs, gloo in the
CPU tests):

    begin   ->  published entries, round 1   (records whose RG:name key was not seen exactly twice locally)
                + copies of the fragment ends whose key lies in another rank's coordinate range
    probe   ->  published entries, round 2   (local couples of names published elsewhere, retracted)
                + pair ends whose key lies in another rank's range; the fragment sort/select start on a
                side stream and overlap the second exchange
    finish  ->  every rank replays the published set in global file order and keeps the pairs it owns;
                pair sort/select; marks on records of other ranks
    apply       flag write

The result equals the reference's single-stream `openge dedup --nosplit -v`, not its own
split-by-chromosome mode.  The per-phase device work is behind `ShardEngine`;
`CudaShardEngine` drives the C ABI (oge_gpu_shard_*), the test suite plugs in a numpy model of the
same protocol (tests/sharded_model.py) to run the orchestration under gloo without a GPU.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

PUB_BYTES, ROUTE_BYTES, MARK_BYTES = 64, 32, 4


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


# --------------------------------------------------------------------------------------- engines
class ShardEngine:
    """Per-rank device work between the exchanges.  Lists are 1-D uint8 torch tensors on `device`."""
    device = "cpu"

    def begin(self): raise NotImplementedError                          # -> (pub, frag_route)
    def probe(self, pub_all, frag_route_all): raise NotImplementedError  # -> (pub2, pair_route)
    def finish(self, pub_both, pair_route_all): raise NotImplementedError  # -> (marks, marks_frag)
    def apply(self, marks_all): raise NotImplementedError
    def flags(self): raise NotImplementedError


class _DevView:
    """Borrowed device memory as a __cuda_array_interface__ object (torch.as_tensor makes a view)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class DevChunk:
    """A list a shard call handed out: device pointer + size, valid until the next call on that context.
    Quacks like the 1-D uint8 tensor the exchanges expect (numel, copy into a tensor slice)."""

    def __init__(self, engine, ptr, nbytes):
        self.engine, self.ptr, self.nbytes = engine, ptr, int(nbytes)

    def numel(self):
        return self.nbytes

    def copy_into(self, dst):      # dst: uint8 tensor slice of the same length on the engine's device
        if self.nbytes:
            self.engine.ctx.copy_d2d(dst.data_ptr(), self.ptr, self.nbytes)

    def tensor(self):
        torch = self.engine.torch
        if not self.nbytes:
            return torch.empty(0, dtype=torch.uint8, device=self.engine.device)
        with torch.cuda.device(self.engine.device):
            return torch.as_tensor(_DevView(self.ptr, self.nbytes), device=self.engine.device).clone()


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
        self.rank = rank
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

    def _take(self, ptr, count, item):
        return DevChunk(self, ptr, count * item)

    def _give(self, t, item):
        torch = self.torch
        assert t.dtype == torch.uint8 and t.numel() % item == 0
        torch.cuda.current_stream(self.device).synchronize()      # the library works on its own stream
        return (t.data_ptr() if t.numel() else None), t.numel() // item

    def begin(self):
        pub, fr = self.ctx.shard_begin()
        return self._take(*pub, PUB_BYTES), self._take(*fr, ROUTE_BYTES)

    def probe(self, pub_all, frag_route_all):
        pub2, pr = self.ctx.shard_probe(*self._give(pub_all, PUB_BYTES), *self._give(frag_route_all, ROUTE_BYTES))
        return self._take(*pub2, PUB_BYTES), self._take(*pr, ROUTE_BYTES)

    def finish(self, pub_both, pair_route_all):
        m, mf = self.ctx.shard_finish(*self._give(pub_both, PUB_BYTES), *self._give(pair_route_all, ROUTE_BYTES))
        return self._take(*m, MARK_BYTES), self._take(*mf, MARK_BYTES)

    def apply(self, marks_all):
        self.ctx.shard_apply(*self._give(marks_all, MARK_BYTES))

    def flags(self):
        return self.ctx.flags()

    def stats(self):
        return self.ctx.stats()

    def close(self):
        self.ctx.close()


# --------------------------------------------------------------------------------------- exchanges
class LocalExchange:
    """All ranks live in this process (several contexts on one GPU, or the numpy model): every
    rank's lists are simply concatenated in rank order.  `outs` = one tuple of lists per rank."""

    def __init__(self):
        self.bytes_moved = 0

    def __call__(self, outs):
        import torch
        outs = [tuple(as_tensor(t) for t in o) for o in outs]
        k = len(outs[0])
        self.bytes_moved += sum(int(t.numel()) for o in outs for t in o) * max(0, len(outs) - 1)
        return tuple(torch.cat([o[j] for o in outs]) if len(outs) > 1 else outs[0][j] for j in range(k))


class AllToAllExchange:
    """One rank per process under torch.distributed.  Every rank's lists go to every rank with
    all_to_all_single (the path's only collective).  Returns, per list, the ranks' contributions
    concatenated in rank order.

    Two wire protocols.  Sized: the byte counts of the lists first, then one payload with uneven
    splits (two collectives and a host round trip in between).  Framed: ONE collective of fixed-size
    frames [k sizes | payload | padding], the frame capacity taken from the previous call at the same
    place in the step (every rank sees every rank's sizes, so all ranks compute the same capacity); a
    rank whose payload does not fit says so in its header and everybody falls back to the sized
    protocol for that call.  Steady-state runs (same file shape step after step) use one collective
    per exchange."""

    HEADER = 64      # bytes: up to 7 int64 sizes + an overflow flag

    def __init__(self, dist, device, timed=False, framed=True):
        import torch
        self.dist, self.torch, self.device = dist, torch, device
        self.world = dist.get_world_size()
        self.bytes_moved = 0
        self.ms = 0.0
        self.timed = timed and str(device).startswith("cuda")
        self.framed = framed
        self.cap = {}          # call site -> frame payload capacity agreed by all ranks
        self.site = 0
        self.calls = {"framed": 0, "sized": 0}

    def new_step(self):
        self.site = 0

    def _split(self, out, per, recv_offsets):
        torch, W = self.torch, self.world
        res = []
        for j in range(len(per[0])):
            parts = []
            for r in range(W):
                o = recv_offsets[r] + sum(per[r][:j])
                parts.append(out[o: o + per[r][j]])
            res.append(torch.cat(parts) if W > 1 else parts[0])
        return tuple(res)

    def _sized(self, mine, payload):
        torch, dist, W = self.torch, self.dist, self.world
        k = len(mine)
        sizes = torch.tensor([int(t.numel()) for t in mine] * W, dtype=torch.int64, device=self.device)
        sizes_all = torch.empty(W * k, dtype=torch.int64, device=self.device)
        dist.all_to_all_single(sizes_all, sizes)
        per = sizes_all.view(W, k).tolist()                      # per[r][j] = bytes of rank r's list j
        recv = [sum(p) for p in per]
        out = torch.empty(sum(recv), dtype=torch.uint8, device=self.device)
        if sum(recv) or payload.numel():
            send = payload.repeat(W) if payload.numel() else payload
            dist.all_to_all_single(out, send, output_split_sizes=recv, input_split_sizes=[int(payload.numel())] * W)
        offs, pos = [], 0
        for r in range(W):
            offs.append(pos)
            pos += recv[r]
        self.calls["sized"] += 1
        return self._split(out, per, offs), per

    def _framed(self, mine, payload, cap):
        torch, dist, W = self.torch, self.dist, self.world
        k = len(mine)
        frame = self.HEADER + cap
        fits = int(payload.numel()) <= cap
        hdr = torch.zeros(8, dtype=torch.int64, device=self.device)
        hdr[:k] = torch.tensor([int(t.numel()) for t in mine], dtype=torch.int64, device=self.device)
        hdr[7] = 0 if fits else 1
        buf = torch.empty(frame, dtype=torch.uint8, device=self.device)
        buf[: self.HEADER] = hdr.view(torch.uint8)
        if fits and payload.numel():
            buf[self.HEADER: self.HEADER + payload.numel()] = payload
        out = torch.empty(W * frame, dtype=torch.uint8, device=self.device)
        dist.all_to_all_single(out, buf.repeat(W))
        heads = out.view(W, frame)[:, : self.HEADER].contiguous().view(torch.int64).view(W, 8).tolist()
        per = [h[:k] for h in heads]
        if any(h[7] for h in heads):
            return None, per
        self.calls["framed"] += 1
        return self._split(out, per, [r * frame + self.HEADER for r in range(W)]), per

    def __call__(self, outs):
        torch = self.torch
        (mine,) = outs
        mine = tuple(as_tensor(t) for t in mine)
        assert len(mine) <= 7
        if self.timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        payload = torch.cat(list(mine)) if len(mine) > 1 else mine[0]
        site, res = self.site, None
        self.site += 1
        if self.framed and site in self.cap:
            res, per = self._framed(mine, payload, self.cap[site])
        if res is None:
            res, per = self._sized(mine, payload)
        # capacity for the next call at this site: a quarter more than the largest contribution seen now
        biggest = max(sum(p) for p in per)
        self.cap[site] = max(4096, (biggest + biggest // 4 + 255) // 256 * 256)
        if self.timed:
            e1.record()
            e1.synchronize()
            self.ms += e0.elapsed_time(e1)
        self.bytes_moved += int(payload.numel()) * (self.world - 1)
        return res


class GatherExchange:
    """The same delivery (every rank's lists to every rank) as the all-gather it is: one
    all_gather_into_tensor of fixed-size frames [sizes | payload | padding] per exchange, the frame
    capacity agreed from the previous call at the same place in the step; the first call at a site, and
    any call where some rank's payload outgrew the frame, gathers the sizes first and then frames of the
    exact maximum.  No per-peer replication of the send buffer, persistent staging buffers."""

    HEADER = 64

    def __init__(self, dist, device, timed=False):
        import torch
        self.dist, self.torch, self.device = dist, torch, device
        self.world = dist.get_world_size()
        self.bytes_moved = 0
        self.ms = 0.0
        self.timed = timed and str(device).startswith("cuda")
        self.cap, self.site = {}, 0
        self.calls = {"framed": 0, "sized": 0}
        self._send, self._recv = {}, {}

    def new_step(self):
        self.site = 0

    def _buffers(self, site, frame):
        torch = self.torch
        if site not in self._send or self._send[site].numel() != frame:
            self._send[site] = torch.empty(frame, dtype=torch.uint8, device=self.device)
            self._recv[site] = torch.empty(self.world * frame, dtype=torch.uint8, device=self.device)
        return self._send[site], self._recv[site]

    def _gather(self, site, mine, payload, cap, flag):
        torch, dist, W, H = self.torch, self.dist, self.world, self.HEADER
        frame = H + cap
        buf, out = self._buffers(site, frame)
        hdr = torch.tensor([int(t.numel()) for t in mine] + [0] * (7 - len(mine)) + [flag], dtype=torch.int64)
        buf[:H] = hdr.view(torch.uint8).to(self.device, non_blocking=True)
        if not flag:
            pos = H
            for t in mine:      # straight into the frame: no intermediate tensors
                n = int(t.numel())
                if n:
                    if isinstance(t, DevChunk):
                        t.copy_into(buf[pos: pos + n])
                    else:
                        buf[pos: pos + n] = t
                pos += n
        dist.all_gather_into_tensor(out, buf)
        heads = out.view(W, frame)[:, :H].contiguous().view(torch.int64).view(W, 8).cpu().tolist()
        return out, heads, frame

    def __call__(self, outs):
        torch, W, H = self.torch, self.world, self.HEADER
        (mine,) = outs
        k = len(mine)
        assert k <= 7
        if self.timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        total = sum(int(t.numel()) for t in mine)
        payload = None
        site = self.site
        self.site += 1
        out = None
        if site in self.cap:
            cap = self.cap[site]
            fits = total <= cap
            out, heads, frame = self._gather(site, mine, payload, cap, 0 if fits else 1)
            if any(h[7] for h in heads):
                out = None
            else:
                self.calls["framed"] += 1
        if out is None:      # sizes first (a frame without payload), then frames of the exact maximum
            _, heads, _ = self._gather(("sz", site), mine, payload, 0, 1)
            cap = max(256, (max(sum(h[:k]) for h in heads) + 255) // 256 * 256)
            out, heads, frame = self._gather(site, mine, payload, cap, 0)
            self.calls["sized"] += 1
        per = [h[:k] for h in heads]
        res = []
        for j in range(k):
            parts = []
            for r in range(W):
                o = r * frame + H + sum(per[r][:j])
                parts.append(out[o: o + per[r][j]])
            res.append(torch.cat(parts))      # a copy: `out` is reused by the next step
        biggest = max(sum(p) for p in per)
        self.cap[site] = max(4096, (biggest + biggest // 4 + 255) // 256 * 256)
        if self.timed:
            e1.record()
            e1.synchronize()
            self.ms += e0.elapsed_time(e1)
        self.bytes_moved += total * (W - 1)
        return tuple(res)


# --------------------------------------------------------------------------------------- the protocol
def run_phases(engines, exchange):
    """Drive one sharded run over the local `engines` (one per process under torch.distributed, or
    all ranks in-process): three exchanges.  Afterwards every engine's flags() are final."""
    import torch
    if hasattr(exchange, "new_step"):
        exchange.new_step()
    pub_all, frag_route_all = exchange([e.begin() for e in engines])
    pub2_all, pair_route_all = exchange([e.probe(pub_all, frag_route_all) for e in engines])
    w = torch.cat([pub_all, pub2_all]) if pub2_all.numel() else pub_all
    marks_a, marks_b = exchange([e.finish(w, pair_route_all) for e in engines])
    marks_all = torch.cat([marks_a, marks_b]) if marks_b.numel() else marks_a
    for e in engines:
        e.apply(marks_all)
    return {"published": int(w.numel()) // PUB_BYTES, "routed": (int(frag_route_all.numel()) + int(pair_route_all.numel())) // ROUTE_BYTES,
            "marks": int(marks_all.numel()) // MARK_BYTES}


def dedup_in_process(bam, world: int, device: int = 0):
    """All `world` ranks as contexts on ONE GPU (tests): -> (flags of the whole file, info dict)."""
    plan, shards = split_bam(bam, world)
    engines = [CudaShardEngine(rec, off, bam.text, bam.refs, plan, r, device=device) for r, (rec, off) in enumerate(shards)]
    try:
        info = run_phases(engines, LocalExchange())
        flags = np.concatenate([e.flags() for e in engines]) if engines else np.zeros(0, np.uint16)
        info["stats"] = [e.stats() for e in engines]
    finally:
        for e in engines:
            e.close()
    return flags, info


# --------------------------------------------------------------------------------------- bench support
def make_rank_shard(workload, scale, rank, world, pinned=True):
    """Synthetic shard of rank `rank` for the weak-scaling bench: the genome is `world` copies of the
    workload's contig set; rank r draws the workload's reads on its own copy (own seed), and every
    rank draws the same small overlay of pairs whose mates lie on two different ranks' contigs
    (0.5 % of the pairs, 10 % of them duplicates of each other) and merges in the overlay records that
    fall on its contigs.  The concatenation of the shards is one coordinate-sorted file.
    -> (records, offsets, header text, contigs, keepalive)"""
    from . import dedup, synth
    cfg, contigs, rgs = synth.config(workload, scale)
    nc = len(contigs)
    if nc * world > 256:
        raise ValueError("contig table holds 256 entries: %d ranks x %d contigs" % (world, nc))
    all_contigs = [("c%d_%s" % (r, name), ln) for r in range(world) for name, ln in contigs]
    hold = {}

    def alloc(tag):
        def f(nbytes):
            if pinned:
                hold[tag] = dedup.PinnedBuffer(nbytes + 64)
                return hold[tag].array[:nbytes]
            return np.empty(nbytes + 64, dtype=np.uint8)[:nbytes]
        return f

    main = synth.restrict(cfg, all_contigs, rank * nc, (rank + 1) * nc, seed=cfg.seed * 1000 + rank, name_base=rank << 40)
    main.cross_contig_frac = 0.0
    rec, offs = synth.generate(main, records_out=alloc("main") if world == 1 else None)
    if world > 1:
        over = synth.restrict(cfg, all_contigs, 0, nc * world, seed=cfg.seed * 7919 + 17, name_base=1 << 60)
        over.n_templates = max(16, int(cfg.n_templates * world * 0.005))
        over.cross_contig_frac, over.dup_frac = 1.0, 0.10
        over.single_frac = over.mate_unmapped_frac = over.unmapped_pair_frac = over.secondary_frac = over.supplementary_frac = 0.0
        orec, ooffs = synth.generate(over)
        keep = synth.records_on_contigs(orec, ooffs, rank * nc, (rank + 1) * nc)
        rec, offs = synth.merge_sorted(rec, offs, orec, ooffs, keep, records_out=alloc("merged"))
    return rec, offs, synth.header_text(all_contigs, rgs), all_contigs, hold


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

    eng = CudaShardEngine(rec, offs, text, contigs, plan, rank, device=local_rank, pinned_ptr=rec.ctypes.data, profile_events=True)
    ex = (AllToAllExchange if os.environ.get("OGE_EXCHANGE") == "alltoall" else GatherExchange)(dist, dev, timed=True)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step():
        # every phase ends with a host-side sync on the library's streams, so events on torch's stream
        # around the whole step see all of it: device phases, exchanges, and whatever overlaps them
        ex.ms = 0.0
        e0.record()
        info = run_phases([eng], ex)
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
    ms, pass_ms, pass_bytes, pass_launches, launches, ex_ms = [], 0.0, 0, 0, 0, 0.0
    for _ in range(args.steps):
        m, st, info, xm = step()
        ms.append(m)
        ex_ms += xm
        pass_ms += st["ms_sort_pass_kernels"]
        pass_bytes += st["sort_pass_bytes"]
        pass_launches += st["sort_pass_launches"]
        launches += st["launches"]
    dist.barrier()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    sampler.mark_end()
    clocks = sampler.stop()

    # max over ranks of the event-timed step
    t = torch.tensor([float(np.mean(ms)), ex_ms / args.steps, float(st["n_duplicates"]), float(n), wall_ms], dtype=torch.float64, device=dev)
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
        run_phases([eng], ex)
        eng.ctx.flags(flags_pin.array.view(np.uint16)[:n])
        dist.barrier()
        e2e.append(time.perf_counter() - t1)
    e2e_s = torch.tensor([float(np.mean(e2e[1:]))], dtype=torch.float64, device=dev)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = torch.tensor([float(rec.nbytes + offs.nbytes), float(n * 2)], dtype=torch.float64, device=dev)
    dist.all_reduce(h2d, op=dist.ReduceOp.SUM)

    if rank == 0:
        peak, peak_src = benchmod.load_peaks()
        achieved = (pass_bytes / 1e9) / (pass_ms * 1e-3) if pass_ms > 0 else 0.0
        traffic = benchmod.load_traffic()
        line = {
            "metric": metric, "value": total_reads / (ms_per_step * 1e-3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": "%s x %d ranks: every rank holds one coordinate range (a copy of the contig set) of a "
                                   "%d-read file; 0.5%% of the pairs have their mates on two different ranks" % (workload_name, world, total_reads),
                       "reads": total_reads, "reads_rank0": n, "l2": "inputs larger than L2",
                       "wall_ms_per_step": float(mx[4]), "exchange_ms_per_step": float(mx[1]), "device_phase_ms_rank0": st["ms_total"],
                       "timing": "CUDA events around each step (phases sync the library streams before returning), max over ranks",
                       "duplicates_flagged": total_dups,
                       "published_entries": info["published"], "routed_entries": info["routed"], "marks_exchanged": info["marks"],
                       "stage_ms_rank0": {k: st[k] for k in ("ms_endbuild", "ms_join", "ms_sort_pair", "ms_sort_frag", "ms_select", "ms_flags")},
                       "parallelism": "range-sharded x%d, three exchanges of small lists per step (%s)" % (world, type(ex).__name__)},
            "e2e": {"value": total_reads / float(e2e_s[0]), "unit": "reads/s", "h2d_bytes_per_step": int(h2d[0]),
                    "d2h_bytes_per_step": int(h2d[1]), "ms_per_step": float(e2e_s[0]) * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "rs_pass_v2", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "traffic": (traffic["ratio"] * pass_bytes / pass_launches) if traffic and pass_launches else None,
                "traffic_source": ("profiles/roofline_traffic.json: dram bytes / algorithmic bytes = %.3f over the ncu --set full "
                                   "capture of the same kernel at C2 size, applied to this run's bytes per launch" % traffic["ratio"]) if traffic else None, "launches_timed": pass_launches,
                         "avg_launch_ms": pass_ms / pass_launches if pass_launches else None,
                         "algorithmic_bytes_per_launch": pass_bytes / pass_launches if pass_launches else None, "rank": 0},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0
