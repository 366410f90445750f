"""profiles/roofline_traffic.json from one `ncu --set full` capture of a whole step (every kernel of the path):

    ncu --set full --clock-control none -k regex:'endbuild|local_|mate_join|pair_check|rs_|select|flags' -s <first launch of a warm step> -c <launches of one step> \\
        -o gpurun_out/step python tools/bench/lj_ablation.py --product --fused-only --workload C2 --scale 1.0 --steps 1
    ncu -i gpurun_out/step.ncu-rep --page raw --csv > gpurun_out/step_raw.csv
    python tools/bench/roofline_traffic.py gpurun_out/step_raw.csv 50000000 "<how it was captured>" > profiles/roofline_traffic.json

Per kernel (name up to the first '<' or '('): launches in the step, mean DRAM bytes (read + write) and duration per launch."""
import csv
import json
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = rows[0]
    ik, ir, iw, it = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}
    out = {}
    for r in rows[2:]:
        name = re.split(r"[<(]", r[ik])[0].replace("void ", "").replace("oge::", "").strip()
        d = out.setdefault(name, {"launches_per_step": 0, "dram_read": 0.0, "dram_write": 0.0, "us": 0.0})
        d["launches_per_step"] += 1
        d["dram_read"] += float(r[ir]) * scale.get(units[ir], 1.0)
        d["dram_write"] += float(r[iw]) * scale.get(units[iw], 1.0)
        d["us"] += float(r[it]) * tscale.get(units[it], 1.0)
    kernels = {}
    for k, d in out.items():
        m = d["launches_per_step"]
        kernels[k] = {"launches_per_step": m, "dram_bytes_per_launch": (d["dram_read"] + d["dram_write"]) / m,
                      "dram_read_per_launch": d["dram_read"] / m, "dram_write_per_launch": d["dram_write"] / m, "us_per_launch_under_ncu": d["us"] / m}
    print(json.dumps({"workload_reads": int(sys.argv[2]), "source": sys.argv[3] if len(sys.argv) > 3 else "", "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main()
