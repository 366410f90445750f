// BGZF inflate on the device (the part the reference does in util/bgzf_input_stream.cpp:65-142 with one
// zlib call per block behind a 50 ms-polled job queue).  BGZF blocks are independent deflate streams of at most 64 KB.
// The compressed file crosses PCIe (a third to a quarter of the inflated bytes) and the records are born in HBM, where
// the dedup path wants them; nothing is staged through host zlib.  Three ways to decode, selected per process
// (oge_gpu_set_inflate_kernel / OGE_INFLATE_KERNEL=engine|warp|threads):
//
//   engine   Blackwell's hardware DECOMPRESS ENGINE (cuMemBatchDecompressAsync, CUDA 12.8+; the device reports
//       CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE): one operation per BGZF block, raw deflate, any byte alignment of source and
//       destination.  Measured on the B200 with BAM-like blocks (profiles/r2_decompress_engine_probe.jsonl): 129-140 GB/s
//       of inflated bytes for 4 K to 64 K blocks per batch, unchanged next to a 55 GB/s host-to-device copy on another
//       stream, 56 ns of host time per submitted operation -- six times the warp kernel below, and it leaves the SMs
//       free.  The default where the device has it.  Two properties the caller must know: dstNumBytes is a scheduling
//       hint, not a bound (the engine writes whatever the stream holds: the destination keeps engine_inflate_slack()
//       bytes free behind the last block and every block's byte count is checked against ISIZE afterwards), and a
//       stream that is not valid deflate ends the batch with cudaErrorLaunchFailure, which is sticky: the process has
//       lost its CUDA context.  The reference exit(-1)s in the same place ("Zlib inflate failed. Aborting.",
//       bgzf_input_stream.cpp:124-128); a caller that must survive corrupt files picks one of the kernels, which
//       report the block and carry on.
//
// and two kernels over the same decoder (inflate_core.cuh):
//
//   bgzf_inflate_warps    one WARP per block, decode state redundant in all lanes, matches copied by 32 lanes.  Correct and
//       simple, but issue-bound: 31 of 32 lanes repeat the same ~40 instructions per literal (79 % issue slots busy,
//       21-23 GB/s of inflated bytes; profiles/r1_inflate_kernels.txt).
//   bgzf_inflate_threads  one THREAD per block: a warp decodes 32 streams, each lane doing useful work, as a state machine
//       that keeps the lanes converged (inflate_lockstep).  The two hot lookup tables of a stream (9-bit literal/length,
//       7-bit distance: 1.25 KB) sit in shared memory at an odd word stride, so that lanes reading the same index hit
//       different banks; the cold arrays (code lengths, canonical symbol order, counts) are thread-local.  160 streams
//       per SM.  (Its first version ran inflate_block<1> per lane: the lanes drifted apart at the first data-dependent
//       branch and the warp executed one lane at a time -- 4.4-5.2 GB/s.)
// OGE_INFLATE_KERNEL=warp|threads picks one; the default is the one that measured faster.
#include <cuda.h>      // types of the driver entry point only; the library does not link libcuda
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "inflate_core.cuh"
#include "kernels.cuh"
#include "oge_gpu_dedup.h"

namespace oge {

// ---------------------------------------------------------------------------------------------- warp per block
constexpr int INF_WARPS = 8;      // warps per CTA; 8 x 3.9 KB of tables = 31 KB of static shared memory

__global__ void __launch_bounds__(INF_WARPS * 32, 4) bgzf_inflate_warps(BgzfParams P) {
    __shared__ oge_inflate::Tables tables[INF_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const oge_inflate::TablesRef T = oge_inflate::tables_ref(&tables[warp]);
    for (uint64_t b = (uint64_t) blockIdx.x * INF_WARPS + warp; b < P.n_blocks; b += (uint64_t) gridDim.x * INF_WARPS) {
        const uint64_t o0 = P.out_off[b], o1 = P.out_off[b + 1];
        if (o1 == o0) continue;      // the empty end-of-file block
        const int rc = oge_inflate::inflate_block<32, oge_inflate::LIT_BITS, oge_inflate::DIST_BITS>(
            P.comp + P.in_off[b] + 18, P.csize[b] - 26, P.out + o0, (uint32_t) (o1 - o0), T, lane);
        if (rc && lane == 0 && atomicCAS(&P.err[0], 0u, (uint32_t) rc) == 0u) P.err[1] = (uint32_t) (P.block_base + b);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- thread per block
constexpr int INT_THREADS = 160;                      // streams per CTA, one CTA per SM
constexpr int INT_LB = 9, INT_DB = 7;                 // primary table widths
constexpr int INT_HOT_U16 = (1 << INT_LB) + (1 << INT_DB) + 2;      // + 2: odd number of 32-bit words per stream
static_assert(((INT_HOT_U16 / 2) & 1) == 1, "an odd word stride spreads equal indices of the 32 lanes over the 32 banks");

struct ColdTables {      // thread-local: touched when a deflate block starts and on the rare long codes
    uint16_t cl_tab[1 << oge_inflate::CL_BITS];
    uint16_t lit_sym[288], dist_sym[32], cl_sym[20];
    uint16_t lit_cnt[16], dist_cnt[16], cl_cnt[16];
    uint8_t lens[320];
    int32_t status;
};

__global__ void __launch_bounds__(INT_THREADS, 1) bgzf_inflate_threads(BgzfParams P) {
    extern __shared__ uint16_t hot[];      // [INT_THREADS][INT_HOT_U16]
    ColdTables cold;
    oge_inflate::TablesRef T;
    T.lit_tab = hot + (size_t) threadIdx.x * INT_HOT_U16;
    T.dist_tab = T.lit_tab + (1 << INT_LB);
    T.cl_tab = cold.cl_tab;
    T.lit_sym = cold.lit_sym; T.dist_sym = cold.dist_sym; T.cl_sym = cold.cl_sym;
    T.lit_cnt = cold.lit_cnt; T.dist_cnt = cold.dist_cnt; T.cl_cnt = cold.cl_cnt;
    T.lens = cold.lens;
    T.status = &cold.status;
    const uint64_t stride = (uint64_t) gridDim.x * INT_THREADS;
    // consecutive lanes take consecutive blocks: similar sizes, so the 32 streams of a warp finish close together.
    // The trip count is warp-uniform (lanes past the end keep the others company inside inflate_lockstep).
    for (uint64_t b = (uint64_t) blockIdx.x * INT_THREADS + threadIdx.x; __any_sync(0xFFFFFFFFu, b < P.n_blocks); b += stride) {
        uint64_t o0 = 0, o1 = 0;
        if (b < P.n_blocks) {
            o0 = P.out_off[b];
            o1 = P.out_off[b + 1];
        }
        const bool active = o1 > o0;      // not past the end, not the empty end-of-file block
        const int rc = oge_inflate::inflate_lockstep<INT_LB, INT_DB>(active ? P.comp + P.in_off[b] + 18 : nullptr, active ? P.csize[b] - 26 : 0,
                                                                     P.out + o0, (uint32_t) (o1 - o0), T, active);
        if (rc && atomicCAS(&P.err[0], 0u, (uint32_t) rc) == 0u) P.err[1] = (uint32_t) (P.block_base + b);
    }
}

// ---------------------------------------------------------------------------------------------- hardware decompress engine
typedef CUresult (*BatchDecompressFn)(CUmemDecompressParams *, size_t, unsigned int, size_t *, CUstream);
typedef CUresult (*DeviceGetAttributeFn)(int *, CUdevice_attribute, CUdevice);

static void *driver_entry(const char *name) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return fn;
}

// -> bit 0: the device inflates raw deflate in hardware; max bytes of one operation in *max_len
static bool engine_query(int device, int *max_len) {
    static int cached_dev[16], cached_ok[16], cached_len[16], n_cached = 0;      // a handful of devices per process
    static std::mutex guard;      // contexts of several devices may be set up by several host threads at once
    std::lock_guard<std::mutex> lock(guard);
    for (int i = 0; i < n_cached; i++)
        if (cached_dev[i] == device) {
            if (max_len) *max_len = cached_len[i];
            return cached_ok[i] != 0;
        }
    int mask = 0, len = 0;
    DeviceGetAttributeFn get = (DeviceGetAttributeFn) driver_entry("cuDeviceGetAttribute");
    const bool have = get && driver_entry("cuMemBatchDecompressAsync") &&
                      get(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, (CUdevice) device) == CUDA_SUCCESS &&
                      get(&len, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH, (CUdevice) device) == CUDA_SUCCESS &&
                      (mask & CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE) && len >= 65536;
    if (n_cached < 16) {
        cached_dev[n_cached] = device; cached_ok[n_cached] = have; cached_len[n_cached] = len;
        n_cached++;
    }
    if (max_len) *max_len = len;
    return have;
}

uint64_t engine_inflate_slack(int device) {
    int len = 0;
    return engine_query(device, &len) ? (uint64_t) len : 0;
}

int engine_inflate_submit(const uint8_t *d_comp, const uint64_t *h_in_off, const uint32_t *h_csize, const uint32_t *h_isize,
                          const uint64_t *h_out_off, uint8_t *d_out, uint32_t *d_act, uint64_t b0, uint64_t b1, cudaStream_t stream,
                          uint64_t *submitted) {
    static BatchDecompressFn submit = (BatchDecompressFn) driver_entry("cuMemBatchDecompressAsync");
    if (!submit) return fail_msg(OGE_ERR_CUDA, "the driver has no cuMemBatchDecompressAsync (CUDA 12.8+)");
    static thread_local std::vector<CUmemDecompressParams> par;
    par.clear();
    par.reserve(b1 - b0);
    for (uint64_t b = b0; b < b1; b++) {
        if (h_isize[b] == 0) continue;      // the empty end-of-file block: nothing to write, act[b] stays 0
        CUmemDecompressParams q;
        memset(&q, 0, sizeof(q));
        q.srcNumBytes = h_csize[b] - 26;      // between the 18-byte member header and the CRC32 + ISIZE trailer
        q.dstNumBytes = h_isize[b];
        q.dstActBytes = d_act + b;
        q.src = d_comp + h_in_off[b] + 18;
        q.dst = d_out + h_out_off[b];
        q.algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
        par.push_back(q);
    }
    if (par.empty()) return 0;
    // batches of 8192 operations: measured 140 GB/s against 129 for 65536 per call and 68 for 226 K per call
    constexpr size_t PER_CALL = 8192;
    for (size_t i = 0; i < par.size(); i += PER_CALL) {
        size_t bad = (size_t) -1;
        const size_t m = par.size() - i < PER_CALL ? par.size() - i : PER_CALL;
        const CUresult r = submit(par.data() + i, m, 0, &bad, (CUstream) stream);
        if (r != CUDA_SUCCESS) return fail_msg(OGE_ERR_CUDA, "cuMemBatchDecompressAsync failed with %d (operation %lld of the batch)", (int) r, (long long) bad);
    }
    *submitted += par.size();
    return 0;
}

__global__ void bgzf_check_sizes(const uint32_t *act, const uint64_t *out_off, uint64_t n_blocks, uint32_t *err) {
    const uint64_t b = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    if ((uint64_t) act[b] != out_off[b + 1] - out_off[b] && atomicCAS(&err[0], 0u, (uint32_t) oge_inflate::INF_ERR_SHORT) == 0u) err[1] = (uint32_t) b;
}

int launch_bgzf_check_sizes(const uint32_t *act, const uint64_t *out_off, uint64_t n_blocks, uint32_t *err, cudaStream_t stream, uint64_t *launches) {
    if (n_blocks == 0) return 0;
    bgzf_check_sizes<<<(uint32_t) ((n_blocks + 255) / 256), 256, 0, stream>>>(act, out_off, n_blocks, err);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- selection
static int g_inflate_mode = -1;      // INFLATE_*; -1 = not chosen (OGE_INFLATE_KERNEL, else engine where the device has one, else warp)

void set_inflate_kernel(int mode) { g_inflate_mode = mode < 0 || mode > INFLATE_ENGINE ? -1 : mode; }

int inflate_mode_for(int device) {
    int mode = g_inflate_mode;
    if (mode < 0) {
        const char *e = getenv("OGE_INFLATE_KERNEL");
        if (e && !strcmp(e, "threads")) mode = INFLATE_THREADS;
        else if (e && !strcmp(e, "warp")) mode = INFLATE_WARP;
        else mode = INFLATE_ENGINE;
    }
    if (mode == INFLATE_ENGINE && !engine_query(device, nullptr)) mode = INFLATE_WARP;      // no engine on this device / driver
    return mode;
}

int launch_bgzf_inflate(const BgzfParams &P, int mode, int sms, cudaStream_t stream, uint64_t *launches) {
    if (P.n_blocks == 0) return 0;
    if (mode == INFLATE_WARP) {
        int per_sm = 0;      // resident CTAs per SM: one wave, blocks are taken with a grid stride
        if ((cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bgzf_inflate_warps, INF_WARPS * 32, 0) != cudaSuccess || per_sm < 1)) per_sm = 2;
        const uint64_t want = (P.n_blocks + INF_WARPS - 1) / INF_WARPS, cap = (uint64_t) sms * per_sm;
        bgzf_inflate_warps<<<(uint32_t) (want < cap ? want : cap), INF_WARPS * 32, 0, stream>>>(P);
    } else if (mode == INFLATE_THREADS) {
        const size_t smem = (size_t) INT_THREADS * INT_HOT_U16 * sizeof(uint16_t);
        // per device and per call: the attribute belongs to the current device's copy of the function
        OGE_CUDA_TRY(cudaFuncSetAttribute(bgzf_inflate_threads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        const uint64_t want = (P.n_blocks + INT_THREADS - 1) / INT_THREADS;
        bgzf_inflate_threads<<<(uint32_t) (want < (uint64_t) sms ? want : (uint64_t) sms), INT_THREADS, smem, stream>>>(P);
    } else {
        return fail_msg(OGE_ERR_INVALID_ARG, "launch_bgzf_inflate: the engine is driven by engine_inflate_submit");
    }
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace oge
