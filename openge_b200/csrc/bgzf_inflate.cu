// BGZF inflate on the device (the part the reference does in util/bgzf_input_stream.cpp:65-142 with one
// zlib call per block behind a 50 ms-polled job queue).  BGZF blocks are independent deflate streams of at most 64 KB.
// The compressed file crosses PCIe (a third to a quarter of the inflated bytes) and the records are born in HBM, where
// the dedup path wants them; nothing is staged through host zlib.  Two kernels over the same decoder (inflate_core.cuh):
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
#include <stdlib.h>
#include <string.h>

#include "inflate_core.cuh"
#include "kernels.cuh"

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
        if (rc && lane == 0 && atomicCAS(&P.err[0], 0u, (uint32_t) rc) == 0u) P.err[1] = (uint32_t) b;
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
        if (rc && atomicCAS(&P.err[0], 0u, (uint32_t) rc) == 0u) P.err[1] = (uint32_t) b;
    }
}

static int g_inflate_mode = -1;      // 0 threads, 1 warps (default); -1 = not chosen yet (OGE_INFLATE_KERNEL decides)

void set_inflate_kernel(int mode) { g_inflate_mode = mode == 0 ? 0 : 1; }

int launch_bgzf_inflate(const BgzfParams &P, int sms, cudaStream_t stream, uint64_t *launches) {
    if (P.n_blocks == 0) return 0;
    if (g_inflate_mode < 0) {
        const char *e = getenv("OGE_INFLATE_KERNEL");
        g_inflate_mode = e && !strcmp(e, "threads") ? 0 : 1;
    }
    const int mode = g_inflate_mode;
    if (mode == 1) {
        int per_sm = 0;      // resident CTAs per SM: one wave, blocks are taken with a grid stride
        if ((cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bgzf_inflate_warps, INF_WARPS * 32, 0) != cudaSuccess || per_sm < 1)) per_sm = 2;
        const uint64_t want = (P.n_blocks + INF_WARPS - 1) / INF_WARPS, cap = (uint64_t) sms * per_sm;
        bgzf_inflate_warps<<<(uint32_t) (want < cap ? want : cap), INF_WARPS * 32, 0, stream>>>(P);
    } else {
        const size_t smem = (size_t) INT_THREADS * INT_HOT_U16 * sizeof(uint16_t);
        // per device and per call: the attribute belongs to the current device's copy of the function
        OGE_CUDA_TRY(cudaFuncSetAttribute(bgzf_inflate_threads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        const uint64_t want = (P.n_blocks + INT_THREADS - 1) / INT_THREADS;
        bgzf_inflate_threads<<<(uint32_t) (want < (uint64_t) sms ? want : (uint64_t) sms), INT_THREADS, smem, stream>>>(P);
    }
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace oge
