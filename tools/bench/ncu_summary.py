"""Summarise `ncu --page raw --csv` and `--page source --csv` exports: python tools/bench/ncu_summary.py RAW.csv [SRC.csv] [kernel-substring]"""
import csv
import re
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]


def raw(path, only=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if only and only not in r[4]:
            continue
        print("=====", r[4][:70], r[8], r[7])
        for h, u, v in zip(hdr, units, r):
            if any(h.startswith(k) for k in KEYS):
                try:
                    fv = float(v)
                except ValueError:
                    continue
                if "stalled" in h and fv < 0.3:
                    continue
                if h.endswith((".max", ".min")) or ".max." in h or ".min." in h or ".sum.p" in h:
                    continue
                print("   %-85s %-10s %s" % (h, u, v))


def src(path, only=None, top=18):
    rows = list(csv.reader(open(path)))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and len(r) > 6:
            cur["rows"].append(r)
    for k in kern[:1]:
        if only and only not in k["name"]:
            continue
        h = k["hdr"]
        isamp, iex = h.index("# Samples"), h.index("Instructions Executed")
        isec = h.index("L2 Theoretical Sectors Global") if "L2 Theoretical Sectors Global" in h else None
        tot = sum(int(r[isamp] or 0) for r in k["rows"])
        totx = sum(int(r[iex] or 0) for r in k["rows"])
        print("-----", k["name"][:70], "samples", tot, "warp-instr", totx)
        order = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][isamp] or 0))[:top]
        for i in sorted(order):
            r = k["rows"][i]
            print("  #%4d %6.2f%% exec=%9s sect=%10s %s" % (i, 100 * int(r[isamp] or 0) / max(tot, 1), r[iex],
                                                             r[isec] if isec is not None else "", r[1].strip()[:80]))
        if isec is not None:
            print("  -- global memory instructions by theoretical L2 sectors")
            mem = [(int(r[isec] or 0), i) for i, r in enumerate(k["rows"]) if (r[isec] or "0") not in ("0", "")]
            for sct, i in sorted(mem, reverse=True)[:14]:
                print("  #%4d sect=%11d exec=%9s %s" % (i, sct, k["rows"][i][iex], k["rows"][i][1].strip()[:80]))


if __name__ == "__main__":
    only = sys.argv[3] if len(sys.argv) > 3 else None
    raw(sys.argv[1], only)
    if len(sys.argv) > 2 and sys.argv[2] != "-":
        src(sys.argv[2], only)
