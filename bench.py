#!/usr/bin/env python
"""bench.py -- reads/sec of the dedup hot path (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2] [--scale S]

A step = one pass of the whole hot path (end-build, mate join, both radix sorts, select, flag
write) over one synthetic coordinate-sorted paired-end BAM batch.  At N=1 the workload is
config C2 of BASELINE.json (50 M reads, 2x150 bp, ~10 % duplicates); at N>1 the workload is config C5
(800 M reads, 30x WGS shape) range-sharded by coordinate: every rank holds 1/8 of it (100 M reads: weak
scaling, the whole of C5 at N=8), the cuts fall inside contigs, and the ends that straddle ranks are
exchanged with NCCL all-to-alls (openge_b200/sharded.py).

  value     reads/s with the records resident in HBM; device time (CUDA events on the library's
            stream, first kernel -> flags final), max over ranks
  e2e       reads/s through the C ABI from pinned HOST buffers: push (H2D) + run + flags (D2H),
            wall clock bracketed by device syncs
  roofline  per kernel (K1 end-build, K2 join, K3 sort pass, K4 select, K5 flags): SURVEY 8(d) algorithmic bytes over
            the CUDA-event time of its launches, against MEASURED_PEAKS.json's HBM copy peak; the largest stage is the
            line's `kernel`; `whole_path` is the 8(d) model over the whole step
  cpu_baseline  the compiled reference (oracle/_ref, its own threads) or the C oracle port on a
            bounded sample of the same workload, on this box's host cores

--impl reference times the reference's own CPU implementation (oracle/_ref/oge_ref_dedup --mem:
MarkDuplicates::runInternal with records preloaded in RAM) on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "reads/sec dedup (bit-exact flags)"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The timed region of this bench is tens of milliseconds,
    which `nvidia-smi -lms` cannot resolve (its first line arrives after the region has ended), so the samples come from
    NVML in-process on a thread, about one per millisecond; `nvidia-smi` is the fall-back when NVML cannot be loaded.
    start() begins sampling (call it before the warm-up so that the thread is already running); mark_begin()/mark_end()
    bracket the timed region and only samples between them are reported."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.samples = []      # (t, sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.t_begin = self.t_end = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _visible_index(self):
        # NVML enumerates physical devices; CUDA_VISIBLE_DEVICES remaps the ordinals torch / the library use
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "").strip()
        if vis:
            parts = [x.strip() for x in vis.split(",") if x.strip()]
            if self.gpu < len(parts) and parts[self.gpu].isdigit():
                return int(parts[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._nvml = pynvml
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._visible_index()), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self._nvml
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = int(reasons_fn(self._h))
                self.samples.append((time.perf_counter(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.0005)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.t_end is None:
            self.mark_end()
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        lo = self.t_begin if self.t_begin is not None else -1.0
        sm, mx, reasons = [], [], set()
        if self._nvml is not None:
            nv = self._nvml
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            inside = [x for x in self.samples if lo <= x[0] <= self.t_end]
            if not inside and self.samples:      # a region shorter than one poll: the nearest sample
                inside = [min(self.samples, key=lambda x: abs(x[0] - self.t_end))]
            for _, mhz, mask in inside:
                sm.append(mhz)
                for name, b in bits.items():
                    if mask & b:
                        reasons.add(name)
            mx = [self.max_mhz] if self.max_mhz else []
            source = "nvml"
        else:
            for t, ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 8 or not (lo <= t <= self.t_end + 0.05):
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            source = "nvidia-smi"
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": source}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": source}


# --------------------------------------------------------------------------------------- NUMA placement
def bind_to_gpu_numa_node(local_rank):
    """One rank per GPU: run this process on the CPUs NVML names as local to its GPU, BEFORE the pinned host buffers are
    allocated and first touched, so that they land on the GPU's NUMA node (with every rank on node 0 the host-to-device
    copies of eight ranks shared one socket's memory: 21 instead of 53 GB/s per GPU).  -> what was done, for the result line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "").strip()
        idx = local_rank
        if vis:
            parts = [x.strip() for x in vis.split(",") if x.strip()]
            if local_rank < len(parts) and parts[local_rank].isdigit():
                idx = int(parts[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1 and 64 * w + b < ncpu]
        if not cpus:
            return {"bound": False, "why": "NVML names no local CPUs"}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "cpus": len(cpus), "first_cpu": cpus[0], "last_cpu": cpus[-1]}
    except Exception as ex:      # no NVML, no permission: the run goes on unbound
        return {"bound": False, "why": "%s: %s" % (type(ex).__name__, ex)}


# --------------------------------------------------------------------------------------- workload
def make_shard(workload, scale, rank, world, pinned):
    """Synthetic records for this rank -> (records array, offsets, header text, contigs).
    world == 1: the named config.  world > 1: handled by openge_b200.sharded (range shards)."""
    from openge_b200 import synth
    if world > 1:
        from openge_b200 import sharded
        return sharded.make_rank_shard(workload, scale, rank, world, pinned)
    cfg, contigs, rgs = synth.config(workload, scale)
    hold = {}

    def alloc(nbytes):
        if pinned:
            from openge_b200 import dedup
            hold["pin"] = dedup.PinnedBuffer(nbytes + 64)
            return hold["pin"].array[:nbytes]
        return np.empty(nbytes + 64, dtype=np.uint8)[:nbytes]

    rec, offs = synth.generate(cfg, records_out=alloc)
    return rec, offs, synth.header_text(contigs, rgs), contigs, hold


def bench_bgzf(ctx, rec, offs, text, contigs, n, flags_want, flags_pin, steps, before_timing=None):
    """reads/s from a BGZF-compressed BAM in host memory to flags in host memory through the C ABI:
    oge_gpu_dedup_push_bgzf (H2D of the compressed file in pieces, each inflated by the hardware decompress engine under
    the upload of the next) -> oge_gpu_dedup_frame -> run -> flags."""
    import ctypes as C
    from openge_b200 import bamhost, bamio
    head = bamio.serialize_bam_stream(bamio.BamFile(text=text, refs=list(contigs), records=np.zeros(0, np.uint8), offsets=np.zeros(1, np.uint64)))
    level = 1
    t0 = time.perf_counter()
    stream = np.empty(len(head) + rec.nbytes, dtype=np.uint8)
    stream[:len(head)] = np.frombuffer(head, dtype=np.uint8)
    stream[len(head):] = rec
    L = bamhost.lib()
    out, out_n = C.c_void_p(), C.c_size_t()
    bamhost._check(L.oge_bgzf_compress(stream.ctypes.data, stream.nbytes, level, 0, C.byref(out), C.byref(out_n)))
    del stream
    from openge_b200 import dedup as _dedup
    comp_pin = _dedup.PinnedBuffer(out_n.value + 64)      # like the raw-record arm: the host buffer is page-locked
    comp = comp_pin.array[:out_n.value]
    comp[:] = np.ctypeslib.as_array((C.c_uint8 * out_n.value).from_address(out.value))
    L.oge_bam_buffer_free(out)
    t_comp = time.perf_counter() - t0
    try:
        # block table from the block headers (what oge_bam_open_bgzf does for a file)
        in_off, csize, isize = [], [], []
        pos = 0
        while pos < out_n.value:
            bs = int(comp[pos + 16]) + (int(comp[pos + 17]) << 8) + 1
            in_off.append(pos)
            csize.append(bs)
            isize.append(int.from_bytes(comp[pos + bs - 4: pos + bs].tobytes(), "little"))
            pos += bs
        in_off = np.asarray(in_off, dtype=np.uint64)
        csize = np.asarray(csize, dtype=np.uint32)
        isize = np.asarray(isize, dtype=np.uint32)
        if before_timing:
            before_timing()      # the host is quiet again before anything is timed
        times, st, wall = [], None, None
        for _ in range(1 + steps):      # first one is warm-up
            ctx.reset()
            ctx.sync()
            t1 = time.perf_counter()
            ctx.push_bgzf(comp.ctypes.data, comp.nbytes, in_off.ctypes.data, csize.ctypes.data, isize.ctypes.data, len(in_off), len(head), None)
            t2 = time.perf_counter()
            ctx.frame(rec.nbytes)
            t3 = time.perf_counter()
            ctx.run()
            t4 = time.perf_counter()
            ctx.flags(flags_pin.array.view(np.uint16))
            t5 = time.perf_counter()
            times.append(t5 - t1)
            wall = {"push_bgzf": (t2 - t1) * 1e3, "frame": (t3 - t2) * 1e3, "run": (t4 - t3) * 1e3, "flags": (t5 - t4) * 1e3}
            st = ctx.stats()
        assert ctx.n == n and np.array_equal(flags_pin.array.view(np.uint16)[:n], flags_want), "flags from the BGZF path differ"
        secs = float(np.mean(times[1:]))
        # the decoder on its own (untimed extra pass): the whole file in ONE piece, i.e. upload, then inflate
        ctx.reset()
        _dedup.set_bgzf_chunk_bytes(2 ** 64 - 1)
        try:
            ctx.push_bgzf(comp.ctypes.data, comp.nbytes, in_off.ctypes.data, csize.ctypes.data, isize.ctypes.data, len(in_off), len(head), None)
            alone = ctx.stats()
        finally:
            _dedup.set_bgzf_chunk_bytes(0)
        return {"value": n / secs, "unit": "reads/s", "h2d_bytes_per_step": int(comp.nbytes + (0 if st["inflate_mode"] == 2 else in_off.nbytes + csize.nbytes)),
                "d2h_bytes_per_step": int(n * 2), "ms_per_step": secs * 1e3, "bgzf_level": level, "bgzf_bytes": int(comp.nbytes),
                "compress_seconds_untimed": t_comp, "host_buffer": "pinned",
                "decoder": {0: "kernel, thread per block", 1: "kernel, warp per block", 2: "hardware decompress engine (cuMemBatchDecompressAsync)"}[st["inflate_mode"]],
                "pieces": st["inflate_pieces"], "host_wall_ms_last_step": wall, "ms_push_bgzf": st["ms_push_bgzf"], "ms_upload": st["ms_inflate_h2d"],
                "ms_inflate_span_overlapped": st["ms_inflate"], "ms_first_inflate_starts_at": st["ms_inflate_start"],
                "decoder_alone": {"ms_upload": alone["ms_inflate_h2d"], "ms_inflate": alone["ms_inflate"],
                                  "inflate_out_GBps": alone["inflate_bytes_out"] / 1e6 / alone["ms_inflate"] if alone["ms_inflate"] else None},
                "ms_frame": st["ms_frame"], "frame_repairs": st["frame_repairs"], "ms_dedup": st["ms_total"]}
    finally:
        comp = None
        comp_pin.free()



# --------------------------------------------------------------------------------------- roofline
def parse_bytes_per_read(rec, offs, n, sample=4000):
    """SURVEY 8(d) stage A, from the workload's own records: 4 + 32 + l_read_name + 4 * n_cigar_op + l_seq (qualities;
    the packed bases are not needed) + tag bytes up to and including the RG value.  Mean over a sample of records."""
    idx = np.linspace(0, n - 1, num=min(n, sample)).astype(np.int64)
    tot = 0
    for i in idx:
        o, e = int(offs[i]), int(offs[i + 1])
        r = rec[o:e]
        l_name = int(r[12])
        n_cig = int(r[16]) | (int(r[17]) << 8)
        l_seq = int.from_bytes(r[20:24].tobytes(), "little")
        t = 36 + l_name + 4 * n_cig + (l_seq + 1) // 2 + l_seq
        tags = r[t:].tobytes()
        k = tags.find(b"RGZ")
        rg = 0
        if k >= 0:
            z = tags.find(b"\0", k + 3)
            rg = (z + 1) if z >= 0 else len(tags)
        tot += 4 + 32 + l_name + 4 * n_cig + l_seq + rg
    return tot / len(idx)


def load_traffic():
    """dram bytes per launch of the path's kernels from the committed ncu --set full captures (profiles/roofline_traffic.json:
    {kernel: {dram_bytes_per_launch, reads, ...}}), or {}."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def roofline_table(n, a_parse, sts, peak, peak_src):
    """Per-kernel roofline rows + the SURVEY 8(d) whole-path figure from the stats of the timed steps.
    achieved = algorithmic bytes / CUDA-event time of that kernel's launches (profile_events), mean over the steps."""
    def mean(f):
        return float(np.mean([f(st) for st in sts]))
    st = sts[-1]
    n_frag, n_pairs, n_dup = st["n_frag_entries"], st["n_pair_entries"], st["n_duplicates"]
    n_pe = 2 * n_pairs + st["n_join_leftovers"]
    E_f, E_h, E_p = 16, 24, 32      # entry sizes fixed by the contract
    tr = load_traffic()
    traffic = tr.get("kernels", {}) if tr.get("workload_reads") == n else {}      # a capture of another workload says nothing here
    P_f, P_p = st["frag_sort_passes"], st["pair_sort_passes"]
    frag_sorted = mean(lambda s: s["ms_sort_frag"]) > 0.01
    ms_k = {k: mean(lambda s, k=k: s["ms_kernel"][k]) for k in st["ms_kernel"]}
    pass_ms = mean(lambda s: s["ms_sort_pass_kernels"])
    pass_launches = mean(lambda s: s["sort_pass_launches"])
    pass_bytes = mean(lambda s: s["sort_pass_bytes"])
    rows = [
        ("K1 end-build", "endbuild_kernel", ms_k["endbuild"], n * a_parse + n_frag * E_f + n_pe * E_h, 1,
         "A parse %.0f B/read + B emit (16 B per end, 24 B per map-eligible read)" % a_parse),
        ("K2 mate join", "local_match_kernel + local_emit_kernel + mate_join_kernel + pair_check_kernel",
         ms_k["match"] + ms_k["emit"] + ms_k["global_join"] + ms_k["check"], n_pe * 2 * E_h + n_pairs * E_p, 4,
         "C join: 2 x 24 B per map-eligible read + 32 B per pair"),
        ("K3 sort pass", "rs_pass_v2", pass_ms, pass_bytes, max(1.0, pass_launches),
         "one LSD pass: 16 B read + 16 B written per entry (the contract's D stage counts 32-byte pair entries: see whole_path)"),
        ("K4 select", "select_kernel", ms_k["select"], (n_pairs + (n_frag if frag_sorted else 0)) * 16 + 4 * n_dup, 1, "E: 16 B per sorted entry + 4 B per duplicate"),
        ("K5 flags", "flags_kernel", ms_k["flags"], 4 * n, 1, "F: 2 B read + 2 B written per record"),
    ]
    stages = []
    for label, kern, ms, alg, launches, what in rows:
        names = [k.strip() for k in kern.split("+")]
        dram = sum(traffic[k]["dram_bytes_per_launch"] * traffic[k]["launches_per_step"] for k in names if k in traffic) if all(k in traffic for k in names) else None
        stages.append({"stage": label, "kernel": kern, "ms_per_step": ms, "algorithmic_bytes_per_step": float(alg),
                       "achieved": (alg / 1e9) / (ms * 1e-3) if ms > 0 else None,
                       "frac": (alg / 1e9) / (ms * 1e-3) / peak if ms > 0 else None,
                       "launches_per_step": launches, "dram_bytes_per_step_ncu": dram,
                       "dram_over_algorithmic": (dram / alg) if dram and alg else None, "bytes": what})
    whole_alg = (n * a_parse + n_frag * E_f + n_pe * E_h + n_pe * 2 * E_h + n_pairs * E_p +
                 (1 + 2 * P_f) * E_f * n_frag + (1 + 2 * P_p) * E_p * n_pairs + E_f * n_frag + E_p * n_pairs + 4 * n_dup + 4 * n)
    # the key widths the contract's worked example assumes (frag 37 b -> 5 passes, pair 70 b -> 9 passes)
    ms_total = mean(lambda s: s["ms_total"])
    dom = max(stages, key=lambda r: r["ms_per_step"])
    return {"bound": "hbm", "kernel": dom["kernel"] + " (" + dom["stage"] + ", the largest stage of the step)",
            "achieved": dom["achieved"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "peak_source": peak_src,
            "traffic": (dom["dram_bytes_per_step_ncu"] / dom["launches_per_step"]) if dom["dram_bytes_per_step_ncu"] else None,
            "traffic_source": tr.get("source") if dom["dram_bytes_per_step_ncu"] else None,
            "avg_launch_ms": dom["ms_per_step"] / dom["launches_per_step"],
            "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_step"] / dom["launches_per_step"],
            "timing": "CUDA events around every launch of the kernel on the library's stream inside the timed steps (profile_events)",
            "stages": stages,
            "whole_path": {"model": "SURVEY 8(d): A parse + B emit + C join + D sort ((1 + 2P) x entry, entry 16 B frag / 32 B pair, "
                                    "P = the passes this run's keys need) + E select + F flags; the fragment D/E terms are counted even "
                                    "when the reduced fragment pass skips that work",
                           "algorithmic_bytes_per_read": whole_alg / n, "ms_per_step": ms_total,
                           "achieved": (whole_alg / 1e9) / (ms_total * 1e-3), "frac": (whole_alg / 1e9) / (ms_total * 1e-3) / peak,
                           "frac_of_nominal_8TBps": (whole_alg / 1e9) / (ms_total * 1e-3) / 8000.0}}


def oracle_flags(records, offsets, text):
    """The C oracle (test infrastructure) over one file: the checker of the untimed parity passes."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    return oracle.markdup(records, offsets, text)


class OracleCheck(threading.Thread):
    """The C oracle on one workload in a host thread (ctypes releases the GIL), untimed and off the GPU's critical path:
    joined before the result line is printed; the line carries the verdict."""

    def __init__(self, rec, offs, text):
        super().__init__(daemon=True)
        self.rec, self.offs, self.text = rec, offs, text
        self.flags = None
        self.error = None
        self.seconds = None

    def run(self):
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle
            t0 = time.perf_counter()
            self.flags = oracle.markdup(self.rec, self.offs, self.text)
            self.seconds = time.perf_counter() - t0
        except Exception as e:      # reported, not swallowed
            self.error = "%s: %s" % (type(e).__name__, e)

    def verdict(self, got):
        self.join()
        if self.error:
            return {"checked": False, "error": self.error}
        bad = int((got != self.flags).sum())
        return {"checked": True, "records": int(len(got)), "mismatches": bad, "oracle_seconds_untimed": self.seconds}


def side_configs(args, local_rank):
    """Device-resident figures of the other single-GPU configs (C1, C3, C4 at their stated sizes), each checked record by
    record against the oracle.  They are reported under config, not as bench lines."""
    from openge_b200 import dedup, synth
    out, checks = {}, {}
    for name in ("C1", "C3", "C4"):
        try:
            bam = synth.make(name, 1.0)
            with dedup.context_for(bam, device=local_rank) as ctx:
                ctx.push(bam.records, bam.offsets)
                ms = []
                for i in range(2 + 3):
                    ctx.run()
                    if i >= 2:
                        ms.append(ctx.stats()["ms_total"])
                got = ctx.flags()
                st = ctx.stats()
            m = float(np.mean(ms))
            out[name] = {"reads": bam.n, "ms_per_step": m, "step_ms": [float(x) for x in ms], "reads_per_s": bam.n / (m * 1e-3), "duplicates": int(st["n_duplicates"]),
                         "stage_ms": {k: st[k] for k in ("ms_endbuild", "ms_join", "ms_sort_pair", "ms_sort_frag", "ms_select", "ms_flags")}}
            checks[name] = (OracleCheck(bam.records, bam.offsets, bam.text), got)
        except Exception as ex:
            out[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    # the oracles after everything has been timed, side by side
    for chk, _ in checks.values():
        chk.start()
    for name, (chk, got) in checks.items():
        out[name]["oracle"] = chk.verdict(got)
    return out

def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------- CPU baselines
def cpu_reference_run(bam, threads, reps=1, timeout=600):
    """The compiled reference on `bam` -> (seconds per rep list, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    if oracle._build.REF_BIN and os.path.exists(oracle._build.REF_BIN):
        try:
            r = oracle.ref_time_mem(bam, reps=reps, threads=threads, timeout=timeout)
            return r["seconds"], "reference"
        except Exception as e:      # the reference's pipeline can hang: fall back to the port
            sys.stderr.write("[bench] compiled reference failed (%s); timing the C oracle port instead\n" % e)
    secs = []
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.markdup(bam.records, bam.offsets, bam.text)
        secs.append(time.perf_counter() - t0)
    return secs, "port"


def sample_bam(workload, n_reads):
    from openge_b200 import synth
    cfg, _, _ = synth.config(workload, 1.0)
    full = max(1, int(cfg.n_templates) * 2)
    return synth.make(workload, max(1e-6, n_reads / full))


def run_reference_arm(args):
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # calibrate on a small sample, then size the per-step sample so the whole run takes ~2-3 minutes
    cal = sample_bam(args.workload, 100_000)
    secs, kind = cpu_reference_run(cal, cores, reps=1, timeout=120)
    rate = cal.n / max(secs[0], 1e-3)
    total_steps = args.steps + args.warmup
    target_s = 150.0 / max(1, total_steps)
    n_reads = int(min(4_000_000, max(200_000, rate * target_s)))
    bam = sample_bam(args.workload, n_reads)
    times = []
    for _ in range(total_steps):
        s, kind = cpu_reference_run(bam, cores, reps=1, timeout=600)
        times.append(s[0])
    timed = times[args.warmup:]
    sec = float(np.mean(timed))
    value = bam.n / sec
    sample = "%d reads of workload %s (same generator, scaled) per step; MarkDuplicates::runInternal with records in RAM" % (bam.n, args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample_reads": bam.n},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores if kind == "reference" else 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_name(args):
    names = {"C1": "C1: 1M-read coordinate-sorted paired-end 2x100bp, single library",
             "C2": "C2: 50M-read paired-end 2x150bp, ~10% duplicates, single GPU",
             "C3": "C3: fragment-heavy mixed SE/PE, unmapped mates, soft clips, several libraries",
             "C4": "C4: exome-like, 40% duplicates, dense equal-key runs",
             "C5": "C5: 800M-read 30x WGS shape, range-sharded"}
    s = names.get(args.workload, args.workload)
    if args.scale != 1.0:
        s += " (scale %g)" % args.scale
    return s


# --------------------------------------------------------------------------------------- our arm
def run_ours(args):
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1 (one rank per GPU)")
    from openge_b200 import dedup
    if dedup.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the dedup path has no CPU fallback")
    dist = None
    if world > 1:
        # NCCL prints its version banner to stdout at some debug levels: the bench prints ONE JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    args.numa = numa
    t_gen = time.perf_counter()
    rec, offs, text, contigs, hold = make_shard(args.workload, args.scale, rank, world, pinned=True)
    n = len(offs) - 1
    t_gen = time.perf_counter() - t_gen

    if world > 1:
        from openge_b200 import sharded
        return sharded.bench(args, rank, world, local_rank, rec, offs, text, contigs, METRIC, workload_name(args))

    max_len = max(l for _, l in contigs)
    ctx = dedup.DedupContext(n_ref=len(contigs), max_ref_len=max_len, device=local_rank, profile_events=True,
                             capacity_records=n, capacity_bytes=rec.nbytes, legacy_join=args.legacy_join)
    ctx.set_header(text)
    offs_pin = dedup.PinnedBuffer(offs.nbytes)
    offs_pin.array.view(np.uint64)[:] = offs
    flags_pin = dedup.PinnedBuffer(n * 2)

    def push_all():
        ctx.push_async(rec.ctypes.data, rec.nbytes, offs_pin.ptr, n)

    # ---- the oracle on the same records, in a host thread, untimed: started once the timed regions that share the host with
    # it are over (a run has host round trips, and a busy host shows in its device time), finished before the next one starts
    check = None
    if not args.no_oracle_check:
        check = OracleCheck(rec, offs, text)

    # ---- device-resident timing
    push_all()
    ctx.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        ctx.run()
    ctx.sync()
    dev_ms, sts, launches = [], [], 0
    sampler.mark_begin()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.run()
        st = ctx.stats()
        dev_ms.append(st["ms_total"])
        sts.append(st)
        launches += st["launches"]
    ctx.sync()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    sampler.mark_end()
    clocks = sampler.stop()
    ms_per_step = float(np.mean(dev_ms))
    value = n / (ms_per_step * 1e-3)
    flags_resident = ctx.flags(flags_pin.array.view(np.uint16)).copy()

    # ---- end to end through the C ABI from host buffers
    e2e_steps = max(1, min(args.steps, 3))
    e2e_t = []
    for _ in range(1 + e2e_steps):      # first one is warm-up
        ctx.reset()
        ctx.sync()
        t1 = time.perf_counter()
        push_all()
        ctx.run()
        ctx.flags(flags_pin.array.view(np.uint16))
        e2e_t.append(time.perf_counter() - t1)
    e2e_s = float(np.mean(e2e_t[1:]))
    flags_e2e = flags_pin.array.view(np.uint16).copy()
    assert np.array_equal(flags_e2e, flags_resident), "e2e flags differ from the device-resident run"
    n_dup = int(((flags_e2e & 0x400) != 0).sum())

    # ---- end to end from a BGZF-compressed BAM held in host memory (what the reference's reader starts from): the
    # compressed bytes cross PCIe, the device inflates (decompress engine), frames, dedups; the flags come back
    e2e_bgzf = None
    if check:
        check.start()      # runs under the (untimed) compression of the BGZF arm's input
    if not args.no_bgzf:
        try:
            e2e_bgzf = bench_bgzf(ctx, rec, offs, text, contigs, n, flags_resident, flags_pin, max(1, min(args.steps, 3)),
                                  before_timing=check.join if check else None)
        except Exception as ex:      # an extra measurement must not take the contract line down
            e2e_bgzf = {"error": "%s: %s" % (type(ex).__name__, ex)}
    if check:
        check.join()

    # ---- roofline: per kernel, the largest stage on top, the SURVEY 8(d) whole-path figure next to it
    peak, peak_src = load_peaks()
    roofline = roofline_table(n, parse_bytes_per_read(rec, offs, n), sts, peak, peak_src)

    # ---- the other single-GPU configs at their stated sizes, device-resident, each against the oracle
    others = side_configs(args, local_rank) if not args.no_side_configs else None

    # ---- parity verdict of the timed workload (the oracle thread started before the timing)
    parity = check.verdict(flags_resident) if check else {"checked": False}
    if parity.get("checked") and parity["mismatches"]:
        raise SystemExit("bench.py: %d of %d flag words differ from the oracle" % (parity["mismatches"], n))

    # ---- CPU baseline on a bounded sample (rank 0, N=1)
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample_n = 2_000_000
        sb = sample_bam(args.workload, sample_n)
        secs, kind = cpu_reference_run(sb, cores, reps=1, timeout=300)
        cpu = {"value": sb.n / secs[0], "unit": "reads/s", "cores": cores if kind == "reference" else 1, "kind": kind,
               "sample": "%d reads of workload %s (same generator, scaled); %s" % (
                   sb.n, args.workload,
                   "compiled reference, MarkDuplicates::runInternal with records in RAM, -t %d" % cores
                   if kind == "reference" else "C oracle port, single thread")}

    line = {
        "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "reads": n, "record_bytes": int(rec.nbytes),
                   "l2": "inputs (%.1f GB) larger than L2" % (rec.nbytes / 1e9), "wall_ms_per_step": wall_ms,
                   "step_ms": [float(x) for x in dev_ms],
                   "duplicates_flagged": n_dup, "gen_seconds": t_gen, "stage_ms": {k: st[k] for k in (
                       "ms_endbuild", "ms_join", "ms_sort_pair", "ms_sort_frag", "ms_select", "ms_flags")},
                   "key_bits": [st["frag_key_bits"], st["pair_key_bits"]],
                   "sort_passes": [st["frag_sort_passes"], st["pair_sort_passes"]],
                   "join": {"pairs_settled_in_cta": st["n_local_pairs"], "records_to_global_join": st["n_join_leftovers"],
                            "pairs_retracted_by_check": st["n_local_retracted"]},
                   "parity_vs_oracle": parity, "other_configs": others},
        "e2e": {"value": n / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": int(rec.nbytes + offs.nbytes),
                "d2h_bytes_per_step": int(n * 2), "ms_per_step": e2e_s * 1e3},
        "e2e_bgzf": e2e_bgzf,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    ctx.close()
    return 0


def main():
    # the contract is ONE JSON line on stdout: whatever libraries print there (NCCL's version banner, for one)
    # goes to stderr, and only the result line is written to the real stdout
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="default: C2 on one GPU (the configuration the metric is quoted on), C5 range-sharded on several")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-oracle-check", action="store_true", help="skip the untimed record-by-record check against the C oracle")
    ap.add_argument("--no-side-configs", action="store_true", help="skip the device-resident figures of C1 / C3 / C4")
    ap.add_argument("--legacy-join", action="store_true", help="A/B: separate end-build and whole-file hash join instead of the fused form")
    ap.add_argument("--no-bgzf", action="store_true", help="skip the extra end-to-end measurement from a BGZF-compressed BAM in host memory")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "C2" if max(args.gpus, env_int("WORLD_SIZE", 1)) <= 1 else "C5"
    if args.warmup < 3 and args.impl == "ours":
        sys.stderr.write("[bench] note: fewer than 3 warm-up steps\n")
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
