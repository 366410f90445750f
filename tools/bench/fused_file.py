"""File-to-file throughput of `openge dedup` on the GPU box: the fused path (oge_dedup_fused) next to the compiled
reference binary, same input file, same flags (--nosplit -v, compression level from --level).

    python tools/bench/fused_file.py --config C2 --scale 0.2 --level 1 [--ref-scale 0.02]

Prints one JSON line: reads/s of both, the fused path's phase split, and whether the two output files are identical
(checked on the reference-sized sample).  The input BAM is written with the host layer's own parallel BGZF writer."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from openge_b200 import _build, bamhost, bamio, synth  # noqa: E402


def write_input(path, name, scale, seed):
    bam = synth.make(name, scale, seed=seed)
    raw = bamio.serialize_bam_stream(bam)
    with open(path, "wb") as f:
        f.write(bamhost.bgzf_compress(raw, 1))
    return bam.n, len(raw)


def run(cmd, timeout):
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, timeout=timeout)
    return time.time() - t0, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--scale", type=float, default=0.2)
    ap.add_argument("--ref-scale", type=float, default=0.02)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--modes", default="", help="comma-separated subset of the fused modes (default: all)")
    a = ap.parse_args()
    fused = _build.ensure_fused()
    ref = _build.REF_BIN if os.path.exists(_build.REF_BIN) else None
    out = {"config": a.config, "level": a.level, "host_threads": a.threads or os.cpu_count()}
    with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
        inp = os.path.join(d, "in.bam")
        n, raw_bytes = write_input(inp, a.config, a.scale, a.seed)
        o1 = os.path.join(d, "fused.bam")
        base = [fused, "dedup", inp, "-o", o1, "-v", "--nopg", "-c", str(a.level)] + (["-t", str(a.threads)] if a.threads else [])
        run(base, 1800)      # warm-up: CUDA context, page cache
        digests = set()
        modes = [("gpu_inflate", []), ("gpu_inflate_pinned", ["--pinned"]), ("cpu_inflate", ["--cpu-inflate"]),
                 ("cpu_inflate_pinned", ["--cpu-inflate", "--pinned"]), ("gpu_inflate_rawbam_out", ["-F", "rawbam"]),
                 ("gpu_deflate", ["--gpu-deflate"]), ("gpu_deflate_pinned", ["--gpu-deflate", "--pinned"])]
        if a.modes:
            modes = [m for m in modes if m[0] in a.modes.split(",")]
        for mode, extra in modes:
            secs, r = run(base + extra, 1800)
            assert r.returncode == 0, r.stderr.decode()[-2000:]
            timing = [l for l in r.stderr.decode().splitlines() if l.startswith("Timing:") or l.startswith("gpu deflate:")]
            out["fused" if mode == "gpu_inflate" else "fused_" + mode] = {
                "reads": n, "raw_bytes": raw_bytes, "file_bytes": os.path.getsize(inp), "out_file_bytes": os.path.getsize(o1), "seconds": secs,
                "reads_per_s": n / secs, "phases": timing if timing else None}
            if "rawbam" not in mode and "deflate" not in mode:
                digests.add(hashlib.sha256(open(o1, "rb").read()).hexdigest())
            if "deflate" in mode:      # identical after decompression, not byte for byte
                out["fused_" + mode]["inflated_sha256"] = hashlib.sha256(bamhost.bgzf_decompress(open(o1, "rb").read())).hexdigest()
        out["all_modes_same_output"] = len(digests) <= 1
        if "fused" in out and any("deflate" in m for m, _ in modes):
            run(base, 1800)
            out["byte_identical_path_inflated_sha256"] = hashlib.sha256(bamhost.bgzf_decompress(open(o1, "rb").read())).hexdigest()
        if ref:
            inp2 = os.path.join(d, "in2.bam")
            n2, _ = write_input(inp2, a.config, a.ref_scale, a.seed)
            o2, o3 = os.path.join(d, "ref.bam"), os.path.join(d, "fused2.bam")
            secs2, r2 = run([ref, "-T", d, "--nosplit", "-v", "-c", str(a.level), inp2, o2] + (["-t", str(a.threads)] if a.threads else []), 1800)
            assert r2.returncode == 0, r2.stderr.decode()[-2000:]
            _, r3 = run([fused, "dedup", inp2, "-o", o3, "--nopg", "-c", str(a.level)], 1800)
            assert r3.returncode == 0
            same = hashlib.sha256(open(o2, "rb").read()).digest() == hashlib.sha256(open(o3, "rb").read()).digest()
            out["reference"] = {"reads": n2, "seconds": secs2, "reads_per_s": n2 / secs2, "output_files_identical": same}
            if "fused" in out:
                out["speedup_file_to_file"] = out["fused"]["reads_per_s"] / out["reference"]["reads_per_s"]
            if "fused_gpu_deflate" in out:
                out["speedup_file_to_file_gpu_deflate"] = out["fused_gpu_deflate"]["reads_per_s"] / out["reference"]["reads_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
