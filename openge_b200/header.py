"""Host-side header helpers for the dedup path: ``@RG ID -> LB -> libraryId``.

Mirrors the reference's resolution exactly (under /root/reference/openge/src):
  * header text is split on newlines; a last line with no trailing newline is dropped
    (util/bam_header.cpp:107-116)
  * ``@RG`` fields: the last ``ID`` / ``LB`` on the line wins (util/bam_header.cpp:81-106)
  * lookup by ID returns the FIRST ``@RG`` with that ID (util/bam_header.h:214-241)
  * no RG tag / unknown ID / empty LB -> "Unknown Library"
    (algorithms/mark_duplicates.cpp:301-318)
  * library ids are dense small integers per distinct library NAME, only equality matters
    (algorithms/mark_duplicates.cpp:282-294)
"""
from __future__ import annotations

UNKNOWN_LIBRARY = "Unknown Library"


def header_lines(text: str):
    lines = text.split("\n")
    # getline() + !in.good(): the final unterminated line is never seen by the reference
    return lines[:-1]


def parse_read_groups(text: str):
    """-> list of (ID, LB) in header order (LB '' when absent)."""
    out = []
    for line in header_lines(text):
        if not line.startswith("@RG\t"):
            continue
        rid, lb = "", ""
        for seg in line[4:].split("\t"):
            tag, data = seg[:2], seg[3:]
            if tag == "ID":
                rid = data
            elif tag == "LB":
                lb = data
        out.append((rid, lb))
    return out


def parse_sequences(text: str):
    """-> list of (SN, LN) from the header text."""
    out = []
    for line in header_lines(text):
        if not line.startswith("@SQ\t"):
            continue
        name, ln = "", -1
        for seg in line[4:].split("\t"):
            tag, data = seg[:2], seg[3:]
            if tag == "SN":
                name = data
            elif tag == "LN":
                try:
                    ln = int(data)
                except ValueError:
                    ln = 0
        out.append((name, ln))
    return out


def library_table(text: str):
    """-> (rg_ids: list[bytes], lib_ids: list[int], unknown_lib_id: int, n_libs: int)

    ``rg_ids`` are unique (first @RG with an ID wins); ``lib_ids[i]`` is the 1-based id of
    that read group's library name; ``unknown_lib_id`` is the id of "Unknown Library".
    """
    names = {}

    def lib_id(name):
        if name not in names:
            names[name] = len(names) + 1
        return names[name]

    unknown = lib_id(UNKNOWN_LIBRARY)
    rg_ids, lib_ids, seen = [], [], set()
    for rid, lb in parse_read_groups(text):
        if rid in seen:
            continue
        seen.add(rid)
        rg_ids.append(rid.encode("latin-1"))
        lib_ids.append(lib_id(lb) if lb else unknown)
    return rg_ids, lib_ids, unknown, len(names)
