// BGZF inflate on the device (the part the reference does in util/bgzf_input_stream.cpp:65-142 with one
// zlib call per block behind a 50 ms-polled job queue).  BGZF blocks are independent deflate streams of at most 64 KB,
// so a file is inflated by one warp per block: the decoder (inflate_core.cuh) keeps its bit-reader state redundantly in
// all 32 lanes, reads compressed words and table entries as warp broadcasts, and copies matches with all lanes.  The
// compressed file crosses PCIe (a third to a quarter of the inflated bytes) and the records are born in HBM, where
// the dedup path wants them; nothing is staged through host zlib.
#include "inflate_core.cuh"
#include "kernels.cuh"

namespace oge {

constexpr int INF_WARPS = 8;      // warps per CTA; 8 x 3.9 KB of tables = 31 KB of static shared memory

__global__ void __launch_bounds__(INF_WARPS * 32, 4) bgzf_inflate_kernel(BgzfParams P) {
    __shared__ oge_inflate::Tables tables[INF_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint64_t b = (uint64_t) blockIdx.x * INF_WARPS + warp; b < P.n_blocks; b += (uint64_t) gridDim.x * INF_WARPS) {
        const uint64_t o0 = P.out_off[b], o1 = P.out_off[b + 1];
        if (o1 == o0) continue;      // the empty end-of-file block
        const int rc = oge_inflate::inflate_block(P.comp + P.in_off[b] + 18, P.csize[b] - 26, P.out + o0, (uint32_t) (o1 - o0), &tables[warp], lane);
        if (rc && lane == 0 && atomicCAS(&P.err[0], 0u, (uint32_t) rc) == 0u) P.err[1] = (uint32_t) b;
        __syncwarp();
    }
}

int launch_bgzf_inflate(const BgzfParams &P, int sms, cudaStream_t stream, uint64_t *launches) {
    if (P.n_blocks == 0) return 0;
    const uint64_t want = (P.n_blocks + INF_WARPS - 1) / INF_WARPS;
    static int per_sm = 0;      // resident CTAs per SM: one wave, blocks are taken with a grid stride
    if (!per_sm) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bgzf_inflate_kernel, INF_WARPS * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    }
    const uint64_t cap = (uint64_t) sms * per_sm;
    bgzf_inflate_kernel<<<(uint32_t) (want < cap ? want : cap), INF_WARPS * 32, 0, stream>>>(P);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace oge
