// Record framing on the device: BamDeserializer::read's chain walk (reference util/bam_deserializer.h:144-172) over the
// inflated stream in HBM -> offsets[].  The chain is sequential by nature (a record's block_size says where the next one
// starts), so it is walked speculatively in parallel and then PROVEN:
//   frame_guess    one warp per 64 KB chunk finds the first position at or after the chunk start from which a chain of
//                  plausible records begins (32 lanes test 32 candidate positions at a time);
//   frame_walk     one thread per chunk walks from the chunk's entry to the first record start beyond its end: record
//                  count and exit position (second call: writes the offsets at the chunk's base rank);
//   verification   chunk 0 starts at byte 0, which is exact.  If exit[k-1] == entry[k] for every k, every chunk was
//                  walked from a true record start, by induction, and the result IS the sequential chain -- no heuristic
//                  is trusted.  Where a guess was wrong the host re-walks that chunk from the proven exit of its
//                  predecessor (frame_walk on one chunk) and carries on; a chain that breaks (block_size outside
//                  [32, 10000], truncated last record) is reported with the reference's message.
#include "kernels.cuh"

namespace oge {

constexpr int FR_THREADS = 128;
constexpr uint32_t FR_MAX_BS = 10000, FR_MIN_BS = 32;      // util/bam_deserializer.h:160

__device__ __forceinline__ uint32_t fr_u32(const uint8_t *p) {      // unaligned little-endian load
    const uintptr_t a = (uintptr_t) p;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t) 3);
    const uint32_t sh = (uint32_t) (a & 3) * 8;
    return sh ? __funnelshift_r(w[0], w[1], sh) : w[0];
}

// Could a record start at `at`?  Only a filter for the guess; nothing downstream trusts it.
__device__ bool fr_plausible(const uint8_t *rec, uint64_t at, uint64_t total, int32_t n_ref) {
    if (at + 36 > total) return false;
    const uint8_t *p = rec + at;
    const uint32_t bs = fr_u32(p);
    if (bs < FR_MIN_BS || bs > FR_MAX_BS || at + 4 + bs > total) return false;
    const int32_t ref = (int32_t) fr_u32(p + 4), pos = (int32_t) fr_u32(p + 8), mref = (int32_t) fr_u32(p + 24), mpos = (int32_t) fr_u32(p + 28);
    if (ref < -1 || ref >= n_ref || mref < -1 || mref >= n_ref || pos < -1 || mpos < -1) return false;
    const uint32_t w3 = fr_u32(p + 12), w4 = fr_u32(p + 16), l_seq = fr_u32(p + 20);
    const uint32_t l_name = w3 & 0xFFu, n_cig = w4 & 0xFFFFu;
    if (l_name == 0 || l_seq > FR_MAX_BS) return false;
    if (32ull + l_name + 4ull * n_cig + ((l_seq + 1) >> 1) + l_seq > bs) return false;
    return p[36 + l_name - 1] == 0;      // the name is NUL-terminated
}

__global__ void __launch_bounds__(FR_THREADS) frame_guess_kernel(FrameParams P) {
    const uint64_t k = (uint64_t) blockIdx.x * (FR_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= P.n_chunks) return;
    if (k == 0) {
        if (lane == 0) P.entry[0] = 0;
        return;
    }
    const uint64_t c0 = k * P.chunk;
    uint64_t found = ~0ull;
    // a record start must exist within 4 + 10000 bytes of any position inside the chain
    for (uint64_t base = c0; base < c0 + FR_MAX_BS + 8 && base < P.total; base += 32) {
        uint64_t at = base + lane;
        bool ok = at < P.total;
        for (int d = 0; ok && d < 3; d++) {      // a chain of three plausible records (or the exact end of the stream)
            if (at == P.total) break;
            ok = fr_plausible(P.rec, at, P.total, P.n_ref);
            if (ok) at += 4ull + fr_u32(P.rec + at);
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
        if (m) {
            found = base + (uint32_t) (__ffs(m) - 1);
            break;
        }
    }
    if (lane == 0) P.entry[k] = found == ~0ull ? c0 : found;      // a wrong guess is caught by the verification
}

// one thread per chunk: walk from entry[k] until the position passes the end of the chunk
__global__ void __launch_bounds__(FR_THREADS) frame_walk_kernel(FrameParams P, uint64_t first_chunk, uint64_t n_walk, int write) {
    const uint64_t j = (uint64_t) blockIdx.x * FR_THREADS + threadIdx.x;
    if (j >= n_walk) return;
    const uint64_t k = first_chunk + j;
    const uint64_t end = (k + 1) * P.chunk < P.total ? (k + 1) * P.chunk : P.total;
    uint64_t at = P.entry[k], cnt = 0;
    uint64_t *out = write ? P.off + P.base[k] : nullptr;
    uint32_t bad = 0;
    while (at < end) {
        if (at + 4 > P.total) { bad = 1; break; }                       // "Expected more bytes reading BAM core"
        const uint32_t bs = fr_u32(P.rec + at);
        if (bs < FR_MIN_BS || bs > FR_MAX_BS) { bad = 2; break; }       // "Invalid BAM block size"
        if (at + 4 + bs > P.total) { bad = 1; break; }
        if (write) out[cnt] = at;
        cnt++;
        at += 4ull + bs;
    }
    if (!write) {
        P.count[k] = cnt;
        P.exit_[k] = at;
        P.bad[k] = bad | (bad == 2 ? (fr_u32(P.rec + at) << 2) : 0u);
    }
    if (write && k == P.n_chunks - 1) P.off[P.base[k] + cnt] = P.total;
}

int launch_frame_guess(const FrameParams &P, cudaStream_t s, uint64_t *launches) {
    if (P.n_chunks == 0) return 0;
    const uint64_t per = FR_THREADS / 32;
    frame_guess_kernel<<<(uint32_t) ((P.n_chunks + per - 1) / per), FR_THREADS, 0, s>>>(P);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_frame_walk(const FrameParams &P, uint64_t first_chunk, uint64_t n_walk, int write, cudaStream_t s, uint64_t *launches) {
    if (n_walk == 0) return 0;
    frame_walk_kernel<<<(uint32_t) ((n_walk + FR_THREADS - 1) / FR_THREADS), FR_THREADS, 0, s>>>(P, first_chunk, n_walk, write);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace oge
