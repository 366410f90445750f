"""Known-answer fixtures of SURVEY.md appendix A.3, rebuilt as raw BAM records.

Expected flags were produced by the compiled reference (`--nosplit -v`) and are re-checked
against it by tests/golden/make_golden.py whenever /root/reference is present.
"""
from __future__ import annotations

import numpy as np

from openge_b200 import bamio

SEQ100 = "ACGT" * 25


def _rec(name, flag, ref, pos1, cigar, mref, mpos1, q, rg="rg1", seq=None):
    seq = SEQ100 if seq is None else seq
    qual = ord(q) - 33 if isinstance(q, str) else q
    tags = bamio.tag_z("RG", rg) if rg else b""
    return bamio.build_record(name, flag, ref, pos1 - 1, 60 if not (flag & 4) else 0, cigar,
                              mref, (mpos1 - 1) if mpos1 else -1, 0, seq, qual, tags)


def fixture1():
    """27 records, two contigs, two libraries.  -> (BamFile, expected dup bit per record)."""
    text = ("@HD\tVN:1.4\tSO:coordinate\n@SQ\tSN:chr1\tLN:100000\n@SQ\tSN:chr2\tLN:100000\n"
            "@RG\tID:rg1\tLB:libA\tSM:s\n@RG\tID:rg2\tLB:libB\tSM:s\n")
    R = [
        # name, flag, ref, pos, cigar, mate ref, mate pos, qual, rg, expected dup
        ("A_tie_first", 99, 0, 1000, "100M", 0, 1300, "I", "rg1", 0),
        ("B_tie_second", 99, 0, 1000, "100M", 0, 1300, "I", "rg1", 1),
        ("H_frag_vs_pairs", 0, 0, 1000, "100M", -1, 0, "I", "rg1", 1),
        ("S_otherlib", 99, 0, 1000, "100M", 0, 1300, "I", "rg2", 0),
        ("A_tie_first", 147, 0, 1300, "100M", 0, 1000, "I", "rg1", 0),
        ("B_tie_second", 147, 0, 1300, "100M", 0, 1000, "I", "rg1", 1),
        ("S_otherlib", 147, 0, 1300, "100M", 0, 1000, "I", "rg2", 0),
        ("D_lowscore", 99, 0, 5000, "100M", 0, 5300, "?", "rg1", 1),
        ("E_highscore", 99, 0, 5000, "100M", 0, 5300, "I", "rg1", 0),
        ("D_lowscore", 147, 0, 5300, "100M", 0, 5000, "?", "rg1", 1),
        ("E_highscore", 147, 0, 5300, "100M", 0, 5000, "I", "rg1", 0),
        ("F_noclip", 99, 0, 8000, "100M", 0, 8300, "I", "rg1", 0),
        ("G_softclip", 99, 0, 8005, "5S95M", 0, 8300, "I", "rg1", 1),
        ("F_noclip", 147, 0, 8300, "100M", 0, 8000, "I", "rg1", 0),
        ("G_softclip", 147, 0, 8300, "100M", 0, 8005, "I", "rg1", 1),
        ("I_fragR_low", 16, 0, 12000, "100M", -1, 0, "5", "rg1", 0),
        ("J_fragR_high", 16, 0, 12010, "90M10S", -1, 0, "I", "rg1", 0),
        ("K_xchrom_first", 99, 0, 20000, "100M", 1, 500, "I", "rg1", 0),
        ("L_xchrom_second", 99, 0, 20000, "100M", 1, 500, "I", "rg1", 1),
        ("M_mateunmapped1", 73, 0, 30000, "100M", 0, 30000, "I", "rg1", 0),
        ("M_mateunmapped1", 133, 0, 30000, "*", 0, 30000, "I", "rg1", 0),
        ("N_mateunmapped2", 73, 0, 30000, "100M", 0, 30000, "I", "rg1", 1),
        ("N_mateunmapped2", 133, 0, 30000, "*", 0, 30000, "I", "rg1", 0),
        ("P_q14", 0, 0, 40000, "100M", -1, 0, "/", "rg1", 1),
        ("R_q15", 0, 0, 40000, "100M", -1, 0, "0", "rg1", 0),
        ("K_xchrom_first", 147, 1, 500, "100M", 0, 20000, "I", "rg1", 0),
        ("L_xchrom_second", 147, 1, 500, "100M", 0, 20000, "I", "rg1", 1),
    ]
    recs = [_rec(n, f, r, p, c, mr, mp, q, rg) for n, f, r, p, c, mr, mp, q, rg, _ in R]
    records, offsets = bamio.concat_records(recs)
    bam = bamio.BamFile(text=text, refs=[("chr1", 100000), ("chr2", 100000)], records=records, offsets=offsets)
    return bam, np.array([x[-1] for x in R], dtype=np.uint8)


def fixture2():
    """Less obvious rules: short wrap, repeated names, supplementary, pre-set 0x400."""
    text = "@HD\tVN:1.4\tSO:coordinate\n@SQ\tSN:chr1\tLN:100000\n@RG\tID:rg1\tLB:libA\tSM:s\n"
    R = [
        # name, flag, pos, cigar, mate pos, seq len, expected flag out
        ("X_long_wraps", 0, 1000, "1000M", 0, 1000, 1024),
        ("Y_short", 0, 1000, "100M", 0, 100, 0),
        ("T_multi", 99, 5000, "100M", 5300, 100, 99),
        ("T_multi", 99, 5000, "100M", 5300, 100, 99),
        ("T_multi", 147, 5300, "100M", 5000, 100, 147),
        ("T_multi", 147, 5300, "100M", 5000, 100, 147),
        ("U_pair", 99, 9000, "100M", 9300, 100, 99),
        ("V_pair", 99, 9000, "100M", 9300, 100, 99),
        ("U_pair", 2147, 9100, "50M50H", 9300, 50, 2147),
        ("U_pair", 147, 9300, "100M", 9000, 100, 147),
        ("V_pair", 147, 9300, "100M", 9000, 100, 147),
        ("W_secondary_predup", 1280, 20000, "100M", 0, 100, 1280),
        ("Z_primary_predup", 1024, 30000, "100M", 0, 100, 0),
    ]
    recs = []
    for n, f, p, c, mp, ls, _ in R:
        seq = ("ACGT" * 250)[:ls]
        recs.append(_rec(n, f, 0, p, c, 0 if mp else -1, mp, "I", "rg1", seq=seq))
    records, offsets = bamio.concat_records(recs)
    bam = bamio.BamFile(text=text, refs=[("chr1", 100000)], records=records, offsets=offsets)
    return bam, np.array([x[-1] for x in R], dtype=np.uint16)


def edge_cases():
    """Extra hand-built cases beyond A.3 (tag walk, RG typing, key ':' ambiguity, negative
    unclipped coordinates, empty names).  Expected values come from the oracle/reference."""
    text = ("@HD\tVN:1.4\tSO:coordinate\n@SQ\tSN:chr1\tLN:100000\n@SQ\tSN:chr2\tLN:5000\n"
            "@RG\tID:a\tLB:L1\n@RG\tID:a:\tLB:L1\n@RG\tID:b\tLB:L2\n@RG\tID:c\n@RG\tID:a\tLB:L9\n")
    recs = []

    def add(name, flag, ref, pos1, cigar, mref, mpos1, q=40, tags=b"", ls=50):
        seq = ("ACGT" * 100)[:ls]
        recs.append(bamio.build_record(name, flag, ref, pos1 - 1, 30, cigar, mref,
                                       (mpos1 - 1) if mpos1 else -1, 0, seq, q, tags))

    # reads whose unclipped start goes negative (leading clip longer than pos)
    add("neg1", 0, 0, 3, "10S40M", -1, 0, tags=bamio.tag_z("RG", "a"))
    add("neg2", 0, 0, 5, "12S38M", -1, 0, tags=bamio.tag_z("RG", "a"))
    # key ambiguity: RG "a:" + ":" + "x"  ==  RG "a" + ":" + ":x" in the reference's map
    add(":x", 99, 0, 100, "50M", 0, 300, tags=bamio.tag_z("RG", "a"))
    add("x", 147, 0, 300, "50M", 0, 100, tags=bamio.tag_z("RG", "a:"))
    add(":x", 99, 0, 100, "50M", 0, 300, tags=bamio.tag_z("RG", "a"))
    add("x", 147, 0, 300, "50M", 0, 100, tags=bamio.tag_z("RG", "a:"))
    # tags in front of RG: integer, string, array; RG found after them
    pre = b"NMC\x03" + b"ASi\x10\x00\x00\x00" + b"XBBs\x02\x00\x00\x00\x01\x00\x02\x00" + bamio.tag_z("MD", "50")
    add("t1", 0, 0, 1000, "50M", -1, 0, tags=pre + bamio.tag_z("RG", "b"))
    add("t2", 0, 0, 1000, "50M", -1, 0, q=20, tags=bamio.tag_z("RG", "b") + pre)
    add("t3", 0, 0, 1000, "50M", -1, 0, q=30, tags=pre)                      # no RG -> Unknown Library
    add("t4", 0, 0, 1000, "50M", -1, 0, q=25, tags=bamio.tag_z("RG", "c"))   # RG without LB -> Unknown Library
    add("t5", 0, 0, 1000, "50M", -1, 0, q=35, tags=bamio.tag_z("RG", "zz"))  # unknown id -> Unknown Library
    # a tag whose type byte is NUL stops the walk before RG is seen
    add("t6", 0, 0, 1000, "50M", -1, 0, q=39, tags=b"XX\x00" + bamio.tag_z("RG", "b"))
    # RG with a non-Z type code is still taken as a string by the reference
    add("t7", 0, 0, 1000, "50M", -1, 0, q=38, tags=b"RGAb\x00")
    # zero-length read name pairs, =/X/N/D/I/P ops, reverse strand ends
    add("", 99, 0, 2000, "10=5X5N10D5I2P25M", 0, 2500, tags=bamio.tag_z("RG", "a"))
    add("", 147, 0, 2500, "20M5S5H", 0, 2000, tags=bamio.tag_z("RG", "a"), ls=25)
    add("dupe", 99, 0, 2000, "20M20N10M20S", 0, 2480, tags=bamio.tag_z("RG", "a"))
    add("dupe", 147, 0, 2480, "5H10S40M", 0, 2000, tags=bamio.tag_z("RG", "a"))
    # paired flag with mate "mapped" but mate refID -1
    add("odd", 65, 0, 3000, "50M", -1, 0, tags=bamio.tag_z("RG", "a"))
    add("odd2", 0, 0, 3000, "50M", -1, 0, q=10, tags=bamio.tag_z("RG", "a"))
    # missing qualities (0xFF) count as 255 each
    add("ff1", 0, 1, 100, "50M", -1, 0, q=255, tags=bamio.tag_z("RG", "a"))
    add("ff2", 0, 1, 100, "50M", -1, 0, q=40, tags=bamio.tag_z("RG", "a"))
    # mapped flag but refID -1; unmapped with coordinates
    add("noref", 0, -1, 0, "50M", -1, 0)
    add("unm", 4, 1, 100, "*", -1, 0)
    records, offsets = bamio.concat_records(recs)
    return bamio.BamFile(text=text, refs=[("chr1", 100000), ("chr2", 5000)], records=records, offsets=offsets)


def sortedness_cases():
    """Small single-end BAMs exercising the "Sorted:" verdict of the reference's Statistics module
    (algorithms/statistics.cpp:89-101), including its quirk: the first record of every contig is exempt
    from the position comparison (last_position is reset to -1 and not set by that record).
    -> dict name -> BamFile"""
    text = ("@HD\tVN:1.4\tSO:unsorted\n@SQ\tSN:chr1\tLN:100000\n@SQ\tSN:chr2\tLN:100000\n@SQ\tSN:chr3\tLN:100000\n"
            "@RG\tID:rg1\tLB:libA\tSM:s\n")
    refs = [("chr1", 100000), ("chr2", 100000), ("chr3", 100000)]

    def mk(rows):
        recs = []
        for k, (ref, pos1, flag) in enumerate(rows):
            if flag & 4:
                recs.append(bamio.build_record("r%03d" % k, flag, ref, pos1 - 1, 0, "*", -1, -1, 0, SEQ100, 30, bamio.tag_z("RG", "rg1")))
            else:
                recs.append(_rec("r%03d" % k, flag, ref, pos1, "100M", -1, 0, "I"))
        records, offsets = bamio.concat_records(recs)
        return bamio.BamFile(text=text, refs=refs, records=records, offsets=offsets)

    cases = {
        "sorted_plain": [(0, 100, 0), (0, 200, 16), (1, 50, 0), (1, 50, 0), (2, 10, 0)],
        "first_of_contig_exempt": [(0, 500, 0), (0, 100, 0), (0, 200, 0), (1, 900, 0), (1, 10, 16), (1, 20, 0)],   # still "Yes"
        "pos_drop_third": [(0, 500, 0), (0, 600, 0), (0, 100, 0)],                                               # "No"
        "contig_drop": [(0, 100, 0), (1, 100, 0), (0, 200, 0)],                                                  # "No"
        "unplaced_skipped": [(0, 100, 0), (0, 300, 0), (-1, 0, 4), (0, 200, 0)],                                 # "No": -1/-1 records are skipped
        "unplaced_between_ok": [(0, 100, 0), (-1, 0, 4), (0, 100, 0), (-1, 0, 4), (-1, 0, 4), (1, 5, 0), (-1, 0, 4), (1, 1, 0), (1, 1, 0)],
        "only_unplaced": [(-1, 0, 4), (-1, 0, 4)],
        "single": [(2, 77, 16)],
    }
    rng = np.random.default_rng(5)
    big = []
    for ref in range(3):      # 3000 sorted records spanning several 1024-record tiles, one violation deep inside
        pos = np.sort(rng.integers(1, 90000, size=1000))
        big += [(ref, int(p), int(rng.integers(0, 2)) * 16) for p in pos]
    cases["big_sorted"] = list(big)
    bad = list(big)
    bad[2049] = (bad[2049][0], 1, 0)
    cases["big_one_violation"] = bad
    edge = list(big)
    edge[1024] = (edge[1024][0], 1, 0)      # first record of the second 1024-tile
    cases["big_violation_at_tile_start"] = edge
    edge2 = list(big)
    edge2[1025] = (edge2[1025][0], 1, 0)    # second record of a tile: its predecessor's predecessor is in the previous tile
    cases["big_violation_second_of_tile"] = edge2
    return {k: mk(v) for k, v in cases.items()}


def header_cases():
    """BAMs whose header text is NOT in the reference's canonical form: what its writer emits for them is defined
    by BamHeader(text).toString() (util/bam_header.cpp:107-262).  -> dict name -> BamFile (records of fixture1)"""
    b1, _ = fixture1()
    sq = "@SQ\tSN:chr1\tLN:100000\n@SQ\tSN:chr2\tLN:100000\n"
    texts = {
        # fields out of the reference's print order, fields it does not know (dropped), KS printed twice
        "reordered_fields": ("@HD\tVN:1.0\tSO:coordinate\tGO:none\n"
                             "@SQ\tUR:file:/x.fa\tSN:chr1\tM5:abc\tLN:100000\tAS:hg0\tSP:human\n@SQ\tSN:chr2\tLN:0100000\n"
                             "@RG\tSM:s\tID:rg1\tLB:libA\tKS:ACGT\tPL:ILLUMINA\tXX:zz\tCN:c\tDS:d\tDT:t\tFO:f\tPG:p\tPI:300\tPU:u\n"
                             "@RG\tID:rg2\tLB:libB\n"
                             "@PG\tVN:1\tID:bwa\tPN:bwa\tCL:bwa mem x\tPP:prev\n@PG\tID:openge\n"
                             "@CO\tfree text\twith a tab\n"),
        # lines in another order: the reference regroups them HD, SQ, RG, PG, CO
        "regrouped_lines": "@CO\tfirst\n@RG\tID:rg1\tLB:libA\n" + sq + "@HD\tVN:1.4\tSO:queryname\n@RG\tID:rg2\tLB:libB\n",
        # no @HD: VN 1.4 / SO unknown is supplied
        "no_hd": sq + "@RG\tID:rg1\tLB:libA\n@RG\tID:rg2\tLB:libB\n",
        # the last line has no newline: the reference drops it
        "unterminated_last_line": "@HD\tVN:1.4\tSO:unsorted\n" + sq + "@RG\tID:rg1\tLB:libA\n@RG\tID:rg2\tLB:libB\n@CO\tdropped",
        # a read group listed twice (the first one wins for the library look-up), one without LB, SO:unknown, an empty @CO
        "repeated_rg_and_no_lb": "@HD\tVN:1.5\tSO:unknown\n" + sq + "@RG\tID:rg1\tLB:libA\n@RG\tID:rg1\tLB:libZ\tSM:x\n@RG\tID:rg2\n@CO\t\n",
        # carriage returns stay inside the last field of a line; numbers are re-printed through atoi; fields may repeat (last one wins)
        "crlf_and_numbers": "@HD\tVN:1.4\tSO:coordinate\r\n".replace("\r", "") + "@SQ\tSN:chr1\tLN:+100000\tAS:a\tAS:b\n@SQ\tSN:chr2\tLN:100000\n"
                            "@RG\tID:rg1\tLB:libA\tPL:x\r\n@RG\tID:rg2\tLB:libB\n@PG\tID:p1\tPN:n\n@PG\tID:p2\tPP:p1\tCL:a b c\n",
    }
    return {k: bamio.BamFile(text=t, refs=list(b1.refs), records=b1.records, offsets=b1.offsets) for k, t in texts.items()}


def shuffled(bam, seed):
    """The same records in a seeded random order (header marked unsorted): input for the coordinate sort."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(bam.n)
    recs = [bam.records[int(bam.offsets[i]):int(bam.offsets[i + 1])].tobytes() for i in perm]
    r, o = bamio.concat_records(recs)
    return bamio.BamFile(text=bam.text.replace("SO:coordinate", "SO:unsorted"), refs=list(bam.refs), records=r, offsets=o)


def name_soup(n=60000, seed=7, pool_div=2.5, n_pos=400, long_names=True, span=2_000_000):
    """Adversarial input for the mate join: names drawn from a small pool, so that a name is seen one to eight times
    anywhere in the file (the reference's map pairs its sightings (1,2), (3,4), ... in FILE order,
    util/picard_structures.h:87-96 + mark_duplicates.cpp:209-246), mates far apart, few distinct positions (long
    duplicate runs), several read groups incl. one the header does not list and reads without RG, names longer than the
    29 bytes a name tag holds that differ only behind them, single-end and mate-unmapped reads in between.  Sorted by
    coordinate; ties keep generation order."""
    rng = np.random.default_rng(seed)
    text = ("@HD\tVN:1.4\tSO:coordinate\n@SQ\tSN:chr1\tLN:%d\n@SQ\tSN:chr2\tLN:%d\n"
            "@RG\tID:rg1\tLB:libA\tSM:s\n@RG\tID:rg2\tLB:libB\tSM:s\n@RG\tID:rg3\tLB:libA\tSM:s\n" % (span + 10000, span + 10000))
    n_names = max(1, int(n / pool_div))
    positions = np.sort(rng.integers(1, span, size=n_pos))
    rows = []
    flags = [99, 147, 83, 163, 65, 129, 97, 145, 73, 0, 16, 1, 1024 | 99]
    rgs = ["rg1", "rg1", "rg1", "rg2", "rg3", "rgX", "rgY", None]
    stem = "instrument:run:flowcell:lane:tile:"          # 34 bytes: longer than a name tag
    for i in range(n):
        k = int(rng.integers(0, n_names))
        name = (stem + "%07d" % k) if (long_names and k % 3 == 0) else "q%d" % k
        flag = flags[int(rng.integers(0, len(flags)))]
        ref = int(rng.integers(0, 2))
        pos = int(positions[int(rng.integers(0, n_pos))])
        mref = ref if rng.random() < 0.8 else 1 - ref
        mpos = int(positions[int(rng.integers(0, n_pos))])
        l = int(rng.integers(20, 37))
        clip = int(rng.integers(0, 4))
        cigar = ("%dS%dM" % (clip, l - clip)) if clip else "%dM" % l
        q = int(rng.integers(10, 41))
        rg = rgs[k % len(rgs)] if rng.random() < 0.97 else rgs[int(rng.integers(0, len(rgs)))]
        rows.append((ref, pos, i, name, flag, cigar, mref, mpos, l, q, rg))
    rows.sort(key=lambda t: (t[0], t[1], t[2]))
    recs = [_rec(nm, f, r, p, c, mr, mp, q, rg, seq=("ACGT" * 10)[:l]) for r, p, _, nm, f, c, mr, mp, l, q, rg in rows]
    records, offsets = bamio.concat_records(recs)
    return bamio.BamFile(text=text, refs=[("chr1", span + 10000), ("chr2", span + 10000)], records=records, offsets=offsets)


def one_name(n=3000, name="*", seed=3):
    """Every record carries the same name (stripped names): the map toggles on one key, pairing records (1,2), (3,4), ..."""
    rng = np.random.default_rng(seed)
    text = "@HD\tVN:1.4\tSO:coordinate\n@SQ\tSN:chr1\tLN:1000000\n@RG\tID:rg1\tLB:libA\tSM:s\n"
    pos = np.sort(rng.integers(1, 5000, size=n))
    recs = [_rec(name, 99 if rng.random() < 0.5 else 147, 0, int(p), "30M", 0, int(pos[int(rng.integers(0, n))]), int(rng.integers(20, 41)),
                 seq=("ACGT" * 10)[:30]) for p in pos]
    records, offsets = bamio.concat_records(recs)
    return bamio.BamFile(text=text, refs=[("chr1", 1000000)], records=records, offsets=offsets)
