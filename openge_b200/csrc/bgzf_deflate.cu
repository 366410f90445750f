// The output side as BGZF members made on the device (output half): stands in for the reference's
// BgzfOutputStream (util/bgzf_output_stream.cpp:59-144: one zlib deflate per 64 KB block on the thread pool, header and
// CRC32 + ISIZE trailer around it) and for the per-record work of its serialiser (the bin BamSerializer recomputes for every
// record, util/bam_serializer.h:88-126) when the caller accepts a file that is identical AFTER DECOMPRESSION rather than
// byte for byte: zlib's match finder is a sequential algorithm, a warp-parallel one finds other (equally valid) matches.
// The byte-identical writer stays on the host (bam_host.cpp, oge_bam_store).
//
// Why on the device: with the inflate on the decompress engine the host's zlib deflate is what a file-to-file run waits
// for (16 host threads: 0.9 GB/s); here the flag-patched records never leave HBM uncompressed -- what
// crosses PCIe back is the finished file.
//
//   fix_bins_kernel      one thread per record: bin = CalculateMinimumBin(pos, pos + reference length of the CIGAR)
//   bgzf_deflate_warps   one warp per block of 65280 bytes (deflate_core.cuh): CRC-32, parse, codes, encode -> a staging slot
//   scan_sizes_kernel    exclusive prefix sum of the member sizes
//   bgzf_assemble_kernel one CTA per block: 18-byte member header, the deflate stream, CRC32 + ISIZE, packed back to back
#include "deflate_core.cuh"
#include "kernels.cuh"
#include "oge_gpu_dedup.h"

namespace oge {

// ---------------------------------------------------------------------------------------------- bins
__device__ __forceinline__ uint32_t minimum_bin(int beg, int end) {      // util/bam_serializer.h:88-98, int arithmetic as there
    --end;
    if ((beg >> 14) == (end >> 14)) return 4681 + (beg >> 14);
    if ((beg >> 17) == (end >> 17)) return 585 + (beg >> 17);
    if ((beg >> 20) == (end >> 20)) return 73 + (beg >> 20);
    if ((beg >> 23) == (end >> 23)) return 9 + (beg >> 23);
    if ((beg >> 26) == (end >> 26)) return 1 + (beg >> 26);
    return 0;
}

__device__ __forceinline__ uint32_t rd32(const uint8_t *p) {      // records are not aligned
    return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24);
}

__global__ void fix_bins_kernel(uint8_t *rec, const uint64_t *off, uint64_t n, uint32_t *err) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t *p = rec + off[i];
    const uint64_t size = off[i + 1] - off[i];
    const uint32_t l_name = p[12], n_cig = (uint32_t) p[16] | ((uint32_t) p[17] << 8);
    if (36ull + l_name + 4ull * n_cig > size) {      // the host writer's check (bam_host.cpp, oge_bam_apply_flags)
        atomicCAS(err, 0u, 1u);
        return;
    }
    const int32_t pos = (int32_t) rd32(p + 8);
    int32_t end = pos;      // bamtools/BamAlignment.cpp:311-350: pos + lengths of M D N = X
    const uint8_t *cg = p + 36 + l_name;
    for (uint32_t k = 0; k < n_cig; k++) {
        const uint32_t c = rd32(cg + 4 * k), op = c & 0xF;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += (int32_t) (c >> 4);
    }
    const uint32_t bin = minimum_bin(pos, end);
    p[14] = (uint8_t) bin;
    p[15] = (uint8_t) (bin >> 8);
}

int launch_fix_bins(uint8_t *rec, const uint64_t *off, uint64_t n, uint32_t *err, cudaStream_t stream, uint64_t *launches) {
    if (n == 0) return 0;
    fix_bins_kernel<<<(uint32_t) ((n + 255) / 256), 256, 0, stream>>>(rec, off, n, err);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- deflate
constexpr int DEF_WARPS = 8;      // warps per CTA: 8 x 9.8 KB of working set + the CRC table = 79 KB, two CTAs per SM

struct DeflateShared {
    oge_deflate::Work work[DEF_WARPS];
    uint32_t crc_table[256];
};

__global__ void __launch_bounds__(DEF_WARPS * 32, 2) bgzf_deflate_warps(DeflateParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    DeflateShared &S = *reinterpret_cast<DeflateShared *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) S.crc_table[i] = oge_deflate::crc_table_entry((uint32_t) i);
    __syncthreads();
    oge_deflate::Work &W = S.work[warp];
    oge_deflate::Seq *seqs = static_cast<oge_deflate::Seq *>(P.seqs) + ((size_t) blockIdx.x * DEF_WARPS + warp) * oge_deflate::MAX_SEQ;
    for (;;) {
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(P.ticket, 1ull);      // blocks are handed out one by one: their cost varies with their content
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        if (b >= P.n_blocks) break;
        const uint64_t at = b * (uint64_t) P.payload;
        const uint32_t n = (uint32_t) (P.total - at < P.payload ? P.total - at : P.payload);
        const uint8_t *src = P.in + at;
        const uint32_t crc = oge_deflate::crc32_block<32>(src, n, S.crc_table, lane);
        const uint32_t bytes = oge_deflate::deflate_block<32>(src, n, P.stage + b * (uint64_t) DEFLATE_STAGE_STRIDE, W, seqs, lane);
        if (lane == 0) {
            P.dsize[b] = bytes;
            P.crc[b] = crc;
        }
        __syncwarp();
    }
}

int deflate_grid(int sms) { return sms * 2; }
size_t deflate_seq_bytes(int sms) { return (size_t) deflate_grid(sms) * DEF_WARPS * oge_deflate::MAX_SEQ * sizeof(oge_deflate::Seq); }

int launch_bgzf_deflate(const DeflateParams &P, int sms, cudaStream_t stream, uint64_t *launches) {
    if (P.n_blocks == 0) return 0;
    if (P.payload == 0 || P.payload > oge_deflate::MAX_BLOCK - 26 - 5 || (P.payload & 3)) return fail_msg(OGE_ERR_INVALID_ARG, "deflate: payload %u", P.payload);
    const size_t smem = sizeof(DeflateShared);
    // per device and per call: the attribute belongs to the current device's copy of the function
    OGE_CUDA_TRY(cudaFuncSetAttribute(bgzf_deflate_warps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    OGE_CUDA_TRY(cudaMemsetAsync(P.ticket, 0, 8, stream));
    const uint64_t want = (P.n_blocks + DEF_WARPS - 1) / DEF_WARPS, cap = (uint64_t) deflate_grid(sms);
    bgzf_deflate_warps<<<(uint32_t) (want < cap ? want : cap), DEF_WARPS * 32, smem, stream>>>(P);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- member offsets
// moff[b] = sum over b' < b of (dsize[b'] + 26); moff[n] = the size of all members.  One CTA, a running carry.
__global__ void __launch_bounds__(1024) scan_sizes_kernel(const uint32_t *dsize, uint64_t n, uint64_t *moff) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = i < n ? (uint64_t) dsize[i] + 26 : 0;
        uint64_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = warp_sum[lane], z = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, z, d);
                if (lane >= d) z += y;
            }
            warp_sum[lane] = z - w;      // exclusive over the warps
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        if (i < n) moff[i] = carry + warp_sum[warp] + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sum[warp] + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) moff[n] = carry_s;
}

// One CTA per member: header (bgzf_output_stream.cpp:117-131 writes the same 18 bytes), stream, CRC32, ISIZE.
__global__ void __launch_bounds__(256) bgzf_assemble_kernel(DeflateParams P, const uint64_t *moff, uint8_t *out) {
    const uint64_t b = blockIdx.x;
    const uint32_t ds = P.dsize[b], csize = ds + 26;
    uint8_t *dst = out + moff[b];
    const uint8_t *src = P.stage + b * (uint64_t) DEFLATE_STAGE_STRIDE;
    const uint64_t at = b * (uint64_t) P.payload;
    const uint32_t n = (uint32_t) (P.total - at < P.payload ? P.total - at : P.payload);
    if (threadIdx.x < 18) {
        const uint8_t head[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, (uint8_t) (csize - 1), (uint8_t) ((csize - 1) >> 8)};
        dst[threadIdx.x] = head[threadIdx.x];
    } else if (threadIdx.x >= 32 && threadIdx.x < 40) {
        const int k = threadIdx.x - 32;
        const uint32_t v = k < 4 ? P.crc[b] : n;
        dst[18 + ds + k] = (uint8_t) (v >> (8 * (k & 3)));
    }
    for (uint32_t i = threadIdx.x; i < ds; i += blockDim.x) dst[18 + i] = src[i];
}

int launch_bgzf_assemble(const DeflateParams &P, uint64_t *moff, uint8_t *out, cudaStream_t stream, uint64_t *launches) {
    if (P.n_blocks == 0) return 0;
    bgzf_assemble_kernel<<<(uint32_t) P.n_blocks, 256, 0, stream>>>(P, moff, out);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_scan_sizes(const uint32_t *dsize, uint64_t n, uint64_t *moff, cudaStream_t stream, uint64_t *launches) {
    scan_sizes_kernel<<<1, 1024, 0, stream>>>(dsize, n, moff);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace oge
