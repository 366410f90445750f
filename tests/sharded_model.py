"""TEST INFRASTRUCTURE: a numpy/python model of the range-sharded protocol (openge_b200/sharded.py,
DESIGN.md section 6) behind the same ShardEngine interface as the CUDA engine.

It exists so that the host orchestration -- ranges, exchanges, phase order -- can run under
torch.distributed/gloo on CPU, and so that the protocol itself is checked against the single-stream
oracle independently of the kernels.  Per-record end fields come from the C oracle
(oracle.markdup(..., want_ends=True), i.e. buildReadEnds of mark_duplicates.cpp:147-164); the pairing,
routing and selection below restate sections A.2.3-A.2.5 of SURVEY.md for one shard.
Pure-python loops: small inputs only.
"""
import numpy as np
import torch

import oracle
from openge_b200.sharded import ShardEngine

PUB = np.dtype([("gidx", "<i8"), ("lib", "<i2"), ("ref", "<i4"), ("coord", "<i4"), ("rev", "u1"), ("paired", "u1"),
                ("score", "<i2"), ("klen", "<i4"), ("key", "S256")])
ROUTE = np.dtype([("kind", "u1"), ("lib", "<i2"), ("ref1", "<i4"), ("coord1", "<i4"), ("orient", "u1"), ("ref2", "<i4"),
                  ("coord2", "<i4"), ("score", "<i2"), ("idx1", "<i8"), ("idx2", "<i8"), ("paired", "u1")])
MARK = np.dtype("<i8")


def _to_t(a):
    return torch.from_numpy(np.frombuffer(a.tobytes(), dtype=np.uint8).copy()) if len(a) else torch.empty(0, dtype=torch.uint8)


def _from_t(t, dt):
    return np.frombuffer(t.numpy().tobytes(), dtype=dt)


def pairing_key(rec, o0, o1):
    """RG value + ':' + read name (mark_duplicates.cpp:210-214); tag walk as BamAlignment.cpp:270-294."""
    p = rec[o0:o1].tobytes()
    l_name, n_cig = p[12], int.from_bytes(p[16:18], "little")
    l_seq = int.from_bytes(p[20:24], "little")
    name = p[36: 36 + max(0, l_name - 1)]
    t = 36 + l_name + 4 * n_cig + (l_seq + 1) // 2 + l_seq
    rg = b""
    sizes = {b"A": 1, b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4}
    while t + 3 <= len(p):
        tag, ty = p[t: t + 2], p[t + 2: t + 3]
        t += 3
        if tag == b"RG":
            e = p.find(b"\0", t)
            rg = p[t: e if e >= 0 else len(p)]
            break
        if ty in sizes:
            t += sizes[ty]
        elif ty in (b"Z", b"H"):
            e = p.find(b"\0", t)
            t = (e if e >= 0 else len(p)) + 1
        elif ty == b"B":
            sub, cnt = p[t: t + 1], int.from_bytes(p[t + 1: t + 5], "little")
            t += 5 + cnt * sizes.get(sub, 1 << 30)
        else:
            break
        if t >= len(p) or p[t] == 0:
            break
    return rg + b":" + name


def _owner_of_key(key: bytes, world: int) -> int:
    """The rank that owns a name: any function of the key bytes all ranks agree on (the CUDA engine uses its key hash)."""
    import zlib
    return zlib.crc32(key) % world


def _bucket(rows, dests, dtype, world, item=None):
    """rows ordered by destination -> Bucketed (stable inside a destination)."""
    from openge_b200.sharded import Bucketed
    order = sorted(range(len(rows)), key=lambda i: dests[i])
    arr = np.array([rows[i] for i in order], dtype=dtype) if rows else np.zeros(0, dtype=dtype)
    counts = [sum(1 for d in dests if d == r) for r in range(world)]
    return Bucketed(_to_t(arr), counts, item or np.dtype(dtype).itemsize)


class ModelShardEngine(ShardEngine):
    device = "cpu"
    entry_bytes = PUB.itemsize

    def __init__(self, records, offsets, header_text, plan, rank):
        self.rank, self.plan = rank, plan
        self.world = plan.world
        self.base = plan.bases[rank]
        self.n = len(offsets) - 1
        self.records, self.offsets = records, offsets
        if self.n:
            _, ends, _ = oracle.markdup(records, offsets, header_text, want_ends=True)
        else:
            ends = np.zeros(0, dtype=oracle.END_DTYPE)
        self.ends = ends
        o = offsets[:-1].astype(np.int64)
        self.flag_in = records[o[:, None] + np.arange(18, 20)].copy().view("<u2").ravel() if self.n else np.zeros(0, "<u2")
        self.dup = np.zeros(self.n, dtype=bool)
        self.splits = list(zip(plan.split_ref, plan.split_pos))

    def key_bytes(self):
        return max([len(pairing_key(self.records, int(self.offsets[i]), int(self.offsets[i + 1]))) for i in range(self.n)], default=0)

    def set_entry_bytes(self, nbytes):
        pass      # the model's entries always hold 256 key bytes

    # rank owning the key range of (ref, coord)
    def owner(self, ref, coord):
        return sum(1 for (r, p) in self.splits if r >= 0 and (r, p) <= (ref, coord))

    def record_owner(self, g):
        return sum(1 for r in range(1, self.world) if self.plan.bases[r] <= g)

    def _pub(self, i):
        e = self.ends[i]
        key = pairing_key(self.records, int(self.offsets[i]), int(self.offsets[i + 1]))
        assert len(key) <= 256
        return (self.base + i, e["lib"], e["ref"], e["coord"], 1 if e["orientation"] == 2 else 0, 1 if e["read2Sequence"] != -1 else 0,
                e["score"], len(key), key)

    @staticmethod
    def _pair(first, second):
        """first = earlier sighting.  Flip rule and orientation of mark_duplicates.cpp:226-243, 169-178."""
        fs, fc, ss, sc = int(first["ref"]), int(first["coord"]), int(second["ref"]), int(second["coord"])
        keep = ss > fs or (ss == fs and sc >= fc)
        a, b = (first, second) if keep else (second, first)
        orient = {(1, 1): 4, (1, 0): 6, (0, 1): 5, (0, 0): 3}[(int(a["rev"]), int(b["rev"]))]
        score = np.array([(int(first["score"]) + int(second["score"])) & 0xFFFF], dtype=np.uint16).view(np.int16)[0]
        return (1, first["lib"], a["ref"], a["coord"], orient, b["ref"], b["coord"], score, a["gidx"], b["gidx"], 1)

    def _frag(self, i):
        e = self.ends[i]
        return (0, e["lib"], e["ref"], e["coord"], int(e["orientation"]), -1, -1, e["score"], self.base + i, -1,
                1 if e["read2Sequence"] != -1 else 0)

    def _pub_bucket(self, ids):
        rows = [self._pub(i) for i in ids]
        return _bucket(rows, [_owner_of_key(r[8], self.world) for r in rows], PUB, self.world)

    def begin(self):
        by_key = {}
        for i in range(self.n):
            if self.ends[i]["pair_eligible"]:
                by_key.setdefault(pairing_key(self.records, int(self.offsets[i]), int(self.offsets[i + 1])), []).append(i)
        self.couples, pub = {}, []
        for key, ids in by_key.items():
            if len(ids) == 2:
                self.couples[key] = ids
            else:
                pub += ids
        self.pairs = []      # ROUTE-shaped tuples owned (so far) by this rank
        # copies of the fragment ends another rank owns leave; the originals stay and are skipped at selection
        self.frags = [self._frag(i) for i in range(self.n) if self.ends[i]["eligible"]]
        out = [f for f in self.frags if self.owner(int(f[2]), int(f[3])) != self.rank]
        # "hashes" of the published keys for every other rank: the model ships the key bytes themselves (8-byte digests)
        import hashlib
        hashes = np.array([np.frombuffer(hashlib.blake2b(self._pub(i)[8], digest_size=8).digest(), dtype="<u8")[0] for i in pub], dtype="<u8")
        return self._pub_bucket(pub), _to_t(hashes), _bucket(out, [self.owner(int(f[2]), int(f[3])) for f in out], ROUTE, self.world)

    def probe(self, hashes_in, frag_route_in):
        import hashlib
        foreign = set(_from_t(hashes_in, "<u8").tolist())
        out = []
        for key in list(self.couples):
            if int(np.frombuffer(hashlib.blake2b(key, digest_size=8).digest(), dtype="<u8")[0]) in foreign:
                out += self.couples.pop(key)
        for r in _from_t(frag_route_in, ROUTE):
            assert self.owner(int(r["ref1"]), int(r["coord1"])) == self.rank
            self.frags.append(tuple(r))
        # the remaining local couples: those whose key another rank owns leave
        route = []
        for key, (i, j) in self.couples.items():
            a = np.array([self._pub(i)], dtype=PUB)[0]
            b = np.array([self._pub(j)], dtype=PUB)[0]
            p = self._pair(a, b)
            (route if self.owner(int(p[2]), int(p[3])) != self.rank else self.pairs).append(p)
        return self._pub_bucket(out), _bucket(route, [self.owner(int(p[2]), int(p[3])) for p in route], ROUTE, self.world)

    def replay(self, pub_in):
        wa = _from_t(pub_in, PUB)
        groups = {}
        seen = set()
        for e in wa:
            assert _owner_of_key(bytes(e["key"])[: e["klen"]], self.world) == self.rank
            if int(e["gidx"]) in seen:
                continue
            seen.add(int(e["gidx"]))
            groups.setdefault(bytes(e["key"])[: e["klen"]], []).append(e)
        route = []
        for key, es in groups.items():
            es.sort(key=lambda e: int(e["gidx"]))
            for k in range(0, len(es) - 1, 2):      # (1,2), (3,4), ...: the toggle of picard_structures.h:87-96
                p = self._pair(es[k], es[k + 1])
                (self.pairs if self.owner(int(p[2]), int(p[3])) == self.rank else route).append(p)
        return _bucket(route, [self.owner(int(p[2]), int(p[3])) for p in route], ROUTE, self.world)

    def _mark(self, g, foreign):
        if self.base <= g < self.base + self.n:
            self.dup[g - self.base] = True
        else:
            foreign.append(g)

    def finish(self, pair_route_in):
        for r in _from_t(pair_route_in, ROUTE):
            assert self.owner(int(r["ref1"]), int(r["coord1"])) == self.rank
            self.pairs.append(tuple(r))
        foreign = []
        groups = {}
        for p in self.pairs:
            groups.setdefault((int(p[1]), int(p[2]), int(p[3]), int(p[4]), int(p[5]), int(p[6])), []).append(p)
        for g in groups.values():      # mark_duplicates.cpp:488-507: best = max score, then smallest idx1
            if len(g) < 2:
                continue
            best = min(g, key=lambda p: (-int(p[7]), int(p[8])))
            for p in g:
                if p is not best:
                    self._mark(int(p[8]), foreign)
                    self._mark(int(p[9]), foreign)
        groups = {}
        for f in self.frags:
            groups.setdefault((int(f[1]), int(f[2]), int(f[3]), int(f[4])), []).append(f)
        for key, g in groups.items():      # :371-390, :515-540
            if len(g) < 2 or all(f[10] for f in g) or self.owner(key[1], key[2]) != self.rank:
                continue
            if any(f[10] for f in g):
                for f in g:
                    if not f[10]:
                        self._mark(int(f[8]), foreign)
            else:
                best = min(g, key=lambda f: (-int(f[7]), int(f[8])))
                for f in g:
                    if f is not best:
                        self._mark(int(f[8]), foreign)
        return _bucket(foreign, [self.record_owner(g) for g in foreign], MARK, self.world)

    def apply(self, marks_in):
        for g in _from_t(marks_in, MARK):
            assert self.base <= g < self.base + self.n
            self.dup[int(g) - self.base] = True

    def flags(self):
        f = self.flag_in.copy()
        prim = (f & 0x100) == 0
        f[prim & self.dup] |= 0x400
        f[prim & ~self.dup] &= np.uint16(~0x400 & 0xFFFF)
        return f
