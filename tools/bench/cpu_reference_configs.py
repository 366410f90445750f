"""The compiled reference (`oracle/_ref/oge_ref_dedup`) timed on the host cores across the BASELINE configs (SURVEY 8(d):
"run C1-C4 in full"; C2 and C5 at a stated slice).  TEST/measurement infrastructure: runs only the reference binary.

    python tools/bench/cpu_reference_configs.py [--quick] > profiles/r1_cpu_reference_configs.json

Two modes per config: `mem` = MarkDuplicates::runInternal alone with the records preloaded in RAM (the hot path);
`file` = `openge dedup --nosplit -v in.bam -o out.bam` at compression level 6, BGZF in and out on /dev/shm (the whole command)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle  # noqa: E402
from openge_b200 import _build, bamhost, bamio, synth  # noqa: E402

FULL = [("C1", 1.0, "1 M reads (full)"), ("C3", 1.0, "10 M records (full)"), ("C4", 1.0, "20 M reads (full)"),
        ("C2", 0.2, "10 M of 50 M reads"), ("C5", 0.0125, "10 M of 800 M reads")]
QUICK = [("C1", 1.0, "1 M reads (full)"), ("C3", 0.1, "1 M of 10 M records"), ("C4", 0.05, "1 M of 20 M reads")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    cores = os.cpu_count()
    out = {"cores": cores, "binary": "oracle/_ref/oge_ref_dedup (the reference's own sources, compiled in place)", "rows": []}
    exe = _build.ensure_ref()
    for name, scale, what in (QUICK if a.quick else FULL):
        bam = synth.make(name, scale)
        row = {"config": name, "scale": scale, "sample": what, "reads": bam.n}
        try:
            r = oracle.ref_time_mem(bam, reps=1, threads=cores, timeout=3600)
            row["mem_seconds"] = r["seconds"][0]
            row["mem_reads_per_s"] = bam.n / r["seconds"][0]
            row["duplicates"] = r["duplicates"]
        except Exception as ex:
            row["mem_error"] = str(ex)[:200]
        with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
            inp, o = os.path.join(d, "in.bam"), os.path.join(d, "out.bam")
            with open(inp, "wb") as f:
                f.write(bamhost.bgzf_compress(bamio.serialize_bam_stream(bam), 6))
            t0 = time.time()
            try:
                p = subprocess.run([exe, "-T", d, "--nosplit", "-v", "-c", "6", "-t", str(cores), inp, o], capture_output=True, timeout=3600)
                row["file_seconds"] = time.time() - t0
                row["file_reads_per_s"] = bam.n / row["file_seconds"]
                row["file_rc"] = p.returncode
            except subprocess.TimeoutExpired:
                row["file_error"] = "did not terminate in 3600 s (the reference's pipeline now and then hangs: SURVEY section 5)"
        out["rows"].append(row)
        print(json.dumps(row), file=sys.stderr)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
