"""In-tree builds of the native pieces (no JIT cache: the .so files travel with the repo).

  libopenge_b200.so   CUDA kernels + C-ABI   (nvcc, sm_100a)      openge_b200/csrc/
  libopenge_b200_testing.so   the same with -DOGE_TESTING (test hooks compiled in; loaded by tests only)
  liboge_bamhost.so   host BAM streaming layer (g++, zlib; no CUDA)   openge_b200/host/bam_host.cpp
  host/_build/oge_dedup_fused   BAM file -> GPU dedup -> BAM file (g++; links both libraries)
  libogesynth.so      synthetic workloads    (gcc)                tools/synth/
  liboge_oracle.so    TEST-ONLY CPU oracle   (gcc)                oracle/
  oracle/_ref/...     the reference itself   (g++, only where /root/reference exists)
"""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "openge_b200", "csrc")
GPU_LIB = os.path.join(ROOT, "openge_b200", "libopenge_b200.so")
SYNTH_LIB = os.path.join(ROOT, "tools", "synth", "libogesynth.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboge_oracle.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "oge_ref_dedup")
HOST_BIN = os.path.join(ROOT, "openge_b200", "host", "_build", "oge_dedup_gpu")
HOST_SORT_BIN = os.path.join(ROOT, "openge_b200", "host", "_build", "oge_mergesort_gpu")
BAMHOST_LIB = os.path.join(ROOT, "openge_b200", "liboge_bamhost.so")
FUSED_BIN = os.path.join(ROOT, "openge_b200", "host", "_build", "oge_dedup_fused")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC", "-shared"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _have(tool):
    from shutil import which
    return which(tool) is not None


def gpu_sources():
    out = []
    for d, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".cu", ".cuh", ".h")):
                out.append(os.path.join(d, f))
    out.append(os.path.join(ROOT, "include", "oge_gpu_dedup.h"))
    return out


class _BuildLock:
    """One builder at a time (several ranks of one torchrun may import the package together)."""

    def __init__(self, path):
        self.path = path

    def __enter__(self):
        import fcntl
        self.f = open(self.path, "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *a):
        import fcntl
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


GPU_TESTING_LIB = os.path.join(ROOT, "openge_b200", "libopenge_b200_testing.so")


def _nvcc_objects(cu, flags, objdir, verbose):
    """One nvcc -c per .cu, all at once; -> object paths (compiler output printed when verbose)."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(objdir, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        return obj, _run(["nvcc"] + flags + inc + ["-c", src, "-o", obj])

    with ThreadPoolExecutor(max_workers=min(len(cu), os.cpu_count() or 4)) as ex:
        res = list(ex.map(one, cu))
    if verbose:
        for _, out in res:
            print(out)
    return [o for o, _ in res]


def build_gpu(force=False, verbose=False, testing=False):
    """libopenge_b200.so, or (testing=True) libopenge_b200_testing.so: the same sources with -DOGE_TESTING, which
    compiles in the test hooks (forced-overflow capacities from the environment, measurement knobs of the sort) that
    the product library does not carry."""
    target = GPU_TESTING_LIB if testing else GPU_LIB
    srcs = gpu_sources()
    cu = [s for s in srcs if s.endswith(".cu")]
    if force or _stale(target, srcs):
        if not _have("nvcc"):
            if os.path.exists(target):
                return target
            raise RuntimeError("nvcc not found and %s is missing" % target)
        with _BuildLock(target + ".lock"):
            if force or _stale(target, srcs):      # somebody else may have built it while we waited
                flags = [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("OGE_NVCC_EXTRA", "").split()
                if testing:
                    flags.append("-DOGE_TESTING=1")
                if verbose:
                    flags += ["-Xptxas", "-v"]
                objs = _nvcc_objects(cu, flags, os.path.join(CSRC, "_obj", "testing" if testing else "product"), verbose)
                tmp = target + ".tmp.%d" % os.getpid()
                _run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", tmp, "-lcudart", "-ldl"])
                os.replace(tmp, target)      # readers never see a half-written library
    return target


def ensure_synth(force=False):
    src = os.path.join(ROOT, "tools", "synth", "oge_synth.c")
    if force or _stale(SYNTH_LIB, [src]):
        _run(["gcc", "-O2", "-fPIC", "-shared", "-pthread", src, "-o", SYNTH_LIB, "-lm"])
    return SYNTH_LIB


def ensure_oracle(force=False):
    """TEST-ONLY: the CPU restatement.  Never called from the product path."""
    src = os.path.join(ROOT, "oracle", "markdup_oracle.c")
    if force or _stale(ORACLE_LIB, [src]):
        _run(["gcc", "-O2", "-fPIC", "-shared", src, "-o", ORACLE_LIB])
    return ORACLE_LIB


def ensure_ref(force=False):
    """TEST-ONLY: compile the reference's dedup path in place (only where its sources exist)."""
    if os.path.isdir("/root/reference/openge/src"):
        d = os.path.join(ROOT, "oracle", "ref_build")
        cli = os.path.join(ROOT, "openge_b200", "host", "refcli", "ref_driver.cpp")      # the command-line front-end shared with the drop-in binaries
        if force or _stale(REF_BIN, [cli, os.path.join(d, "Makefile")]):
            _run(["make", "-C", d] + (["-B"] if force else []))
    return REF_BIN if os.path.exists(REF_BIN) else None


def ensure_host(force=False):
    """The drop-in `openge dedup` binary: the reference's pipeline compiled in place with this repo's
    MarkDuplicates (openge_b200/host/mark_duplicates_gpu.cpp) linked instead of the reference's.
    Needs the reference sources; elsewhere the prebuilt binary (it travels with the snapshot) is used."""
    if os.path.isdir("/root/reference/openge/src"):
        d = os.path.join(ROOT, "openge_b200", "host")
        deps = [os.path.join(d, "mark_duplicates_gpu.cpp"), os.path.join(d, "read_sorter_gpu.cpp"), os.path.join(d, "record_batch.h"),
                os.path.join(d, "Makefile"), GPU_LIB,
                os.path.join(ROOT, "openge_b200", "host", "refcli", "ref_driver.cpp"), os.path.join(ROOT, "include", "oge_gpu_dedup.h")]
        if force or _stale(HOST_BIN, deps) or _stale(HOST_SORT_BIN, deps):
            _run(["make", "-C", d] + (["-B"] if force else []))
    return HOST_BIN if os.path.exists(HOST_BIN) else None


def ensure_bamhost(force=False):
    """The host-side BAM streaming layer (include/oge_bam_host.h): parallel BGZF codec, header model, framing."""
    src = os.path.join(ROOT, "openge_b200", "host", "bam_host.cpp")
    hdr = os.path.join(ROOT, "include", "oge_bam_host.h")
    if force or _stale(BAMHOST_LIB, [src, hdr]):
        with _BuildLock(BAMHOST_LIB + ".lock"):
            if force or _stale(BAMHOST_LIB, [src, hdr]):
                tmp = BAMHOST_LIB + ".tmp.%d" % os.getpid()
                _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "include"), src, "-o", tmp, "-lz"])
                os.replace(tmp, BAMHOST_LIB)
    return BAMHOST_LIB


def ensure_fused(force=False):
    """`openge dedup` as one fused path: oge_bam_load -> libopenge_b200.so -> oge_bam_store.  Needs no reference sources."""
    d = os.path.join(ROOT, "openge_b200", "host")
    srcs = [os.path.join(d, "dedup_fused_main.cpp"), os.path.join(d, "bam_host.cpp")]
    deps = srcs + [os.path.join(ROOT, "include", "oge_bam_host.h"), os.path.join(ROOT, "include", "oge_gpu_dedup.h"), GPU_LIB]
    if force or _stale(FUSED_BIN, deps):
        if not _have("g++"):
            return FUSED_BIN if os.path.exists(FUSED_BIN) else None
        os.makedirs(os.path.dirname(FUSED_BIN), exist_ok=True)
        with _BuildLock(os.path.join(ROOT, "openge_b200", "liboge_fused.lock")):
            if force or _stale(FUSED_BIN, deps):
                tmp = FUSED_BIN + ".tmp.%d" % os.getpid()
                _run(["g++", "-std=c++17", "-O2", "-pthread", "-I", os.path.join(ROOT, "include")] + srcs +
                     ["-o", tmp, "-L", os.path.join(ROOT, "openge_b200"), "-lopenge_b200", "-Wl,-rpath,$ORIGIN/../..", "-lz"])
                os.replace(tmp, FUSED_BIN)
    return FUSED_BIN


def build_all(verbose=False):
    ensure_synth()
    ensure_oracle()
    ensure_ref()
    lib = build_gpu(verbose=verbose)
    build_gpu(verbose=verbose, testing=True)
    ensure_bamhost()
    ensure_fused()
    ensure_host()
    return lib
