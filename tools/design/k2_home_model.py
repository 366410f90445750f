"""Design check for the locality-preserving mate table planned for K2 ("K2, a plan for the next round").

Not product code and not a performance measurement: a numpy model of WHERE the table accesses would land, run on the
synthetic configs, answering the questions that decide whether the scheme is worth building:
  1. do both mates of a pair compute the same home tile from their own record?  (must be all of them on well-formed data;
     the rest would go to the leftover pass, which re-joins by plain hash and keeps the result exact)
  2. how unevenly do the homes load the table (dense equal-coordinate runs of C4)?  -> window load, linear-probe lengths
  3. how many distinct 32-byte slots does a wave of in-flight records touch, and over how wide a stretch of the table,
     against the hash-addressed table of today?

    python tools/design/k2_home_model.py [C2 0.02] [C3 0.1] [C4 0.05]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openge_b200 import synth  # noqa: E402

TILE = 128            # records per tile (K1's tile)
SLOTS_PER_TILE = 128  # table slots reserved per tile of records
W = 1 << 16           # probe window, slots
WAVE = 300_000        # records in flight at once (148 SMs x 2048 threads)


def fields(bam):
    off = bam.offsets[:-1].astype(np.int64)

    def i32(at):
        return bam.records[off[:, None] + np.arange(at, at + 4)].copy().view("<i4").ravel()

    ref, pos, mref, mpos = i32(4), i32(8), i32(24), i32(28)
    flag = bam.records[off[:, None] + np.arange(18, 20)].copy().view("<u2").ravel().astype(np.int64)
    l_name = bam.records[off + 12].astype(np.int64)
    # name hash: FNV over the name bytes (any 64-bit hash will do for the model)
    h = np.full(len(off), 0xcbf29ce484222325, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k in range(int(l_name.max())):
            use = k < l_name - 1
            b = bam.records[np.minimum(off + 36 + k, len(bam.records) - 1)].astype(np.uint64)
            h = np.where(use, (h ^ b) * np.uint64(0x100000001b3), h)
    return ref, pos, mref, mpos, flag, h


def model(name, scale):
    bam = synth.make(name, scale)
    ref, pos, mref, mpos, flag, h = fields(bam)
    n = len(ref)
    # records that enter the mate map (mark_duplicates.cpp:202-209): mapped, primary, paired, mate mapped
    pe = ((flag & 0x4) == 0) & (ref >= 0) & ((flag & 0x100) == 0) & ((flag & 0x1) != 0) & ((flag & 0x8) == 0)
    key = (ref.astype(np.int64) << 32) | (pos.astype(np.int64) & 0xFFFFFFFF)
    mkey = (mref.astype(np.int64) << 32) | (mpos.astype(np.int64) & 0xFFFFFFFF)
    lo = np.minimum(key, mkey)
    # tile starts: key of the first record of every tile, made monotone (unplaced records at the end count as +inf)
    big = np.int64(1) << 62
    skey = np.where(ref >= 0, key, big)
    sorted_input = bool(np.all(np.diff(skey) >= 0))
    starts = np.maximum.accumulate(skey[::TILE])
    home = np.searchsorted(starts, lo, side="right") - 1      # last tile whose first record is <= lo
    home = np.clip(home, 0, len(starts) - 1)
    idx = np.nonzero(pe)[0]
    # 1. mates agree?
    order = np.argsort(h[idx], kind="stable")
    hs, hi = h[idx][order], idx[order]
    same = hs[1:] == hs[:-1]
    a, b = hi[:-1][same], hi[1:][same]
    agree = float(np.mean(home[a] == home[b])) if len(a) else 1.0
    # 2. load per window, probe lengths with linear probing from home*SLOTS_PER_TILE + (hash mod W)
    n_slots = len(starts) * SLOTS_PER_TILE + W
    first = np.ones(len(hi), dtype=bool)
    first[1:] = hs[1:] != hs[:-1]
    names = hi[first]                                   # one representative per name (= one table entry)
    slot0 = home[names] * SLOTS_PER_TILE + (h[names] % np.uint64(W)).astype(np.int64)
    table = np.zeros(n_slots + 65536, dtype=bool)
    probes = np.zeros(len(names), dtype=np.int64)
    for j in np.argsort(names, kind="stable"):          # insertion in file order
        s = int(slot0[j])
        k = 0
        while table[s + k]:
            k += 1
        table[s + k] = True
        probes[j] = k + 1
    nw = n_slots // W
    win_load = table[: nw * W].reshape(nw, W).mean(axis=1) if nw else np.array([table.mean()])
    # 3. distinct slots touched by a wave of records: locality scheme vs hash addressing over a table of n_pe slots
    touched_local, touched_hash, span_local = [], [], []
    hash_slot = (h % np.uint64(max(1, int(pe.sum()) + 1024))).astype(np.int64)
    loc_slot = home * SLOTS_PER_TILE + (h % np.uint64(W)).astype(np.int64)
    for w0 in range(0, n, WAVE):
        sel = idx[(idx >= w0) & (idx < w0 + WAVE)]
        if len(sel) == 0:
            continue
        touched_local.append(len(np.unique(loc_slot[sel])))
        touched_hash.append(len(np.unique(hash_slot[sel])))
        span_local.append(int(loc_slot[sel].max() - loc_slot[sel].min()))
    return {"config": name, "scale": scale, "records": n, "in_mate_map": int(pe.sum()), "names": int(len(names)), "sorted_input": sorted_input,
            "mates_with_equal_home": agree,
            "probe_len_mean": float(probes.mean()), "probe_len_p999": float(np.quantile(probes, 0.999)), "probe_len_max": int(probes.max()),
            "window_load_mean": float(win_load.mean()), "window_load_max": float(win_load.max()),
            "wave_records": WAVE, "wave_table_span_MB_local": float(np.mean(span_local)) * 32 / 1e6,
            "wave_table_span_MB_hash": float(pe.sum() + 1024) * 32 / 1e6,
            "wave_distinct_slots_local": float(np.mean(touched_local)), "wave_distinct_slots_hash": float(np.mean(touched_hash))}


def main():
    args = sys.argv[1:]
    cases = [(args[i], float(args[i + 1])) for i in range(0, len(args), 2)] or [("C2", 0.02), ("C3", 0.1), ("C4", 0.05)]
    for name, scale in cases:
        print(json.dumps(model(name, scale)))


if __name__ == "__main__":
    main()
