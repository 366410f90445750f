// Probe of the Blackwell hardware decompress engine (cuMemBatchDecompressAsync, CUDA 12.8+) on BGZF-shaped input:
// raw deflate streams of <= 64 KB, arbitrary byte alignment of source and destination, hundreds of thousands of streams
// in one batch.  Measurement tool only (not part of the library); build: nvcc -O2 de_probe.cu -o de_probe -lcuda -lz
//   de_probe [n_blocks] [level]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <vector>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_ = 0; cuGetErrorString(r_, &s_); \
    printf("{\"error\": \"%s -> %d %s\", \"line\": %d}\n", #x, (int) r_, s_ ? s_ : "?", __LINE__); return 2; } } while (0)
#define RT(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("{\"error\": \"%s -> %s\", \"line\": %d}\n", #x, cudaGetErrorString(r_), __LINE__); return 2; } } while (0)

static uint64_t rng_state = 88172645463325252ull;
static inline uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t) (rng_state >> 16); }

// something like BAM records: a fixed-ish core, a name with a running number, 4-bit bases (random), qualities from a small alphabet
static void fill_block(uint8_t *p, int n) {
    int i = 0;
    static uint32_t serial = 0;
    while (i < n) {
        int rec = 296;
        if (i + rec > n) rec = n - i;
        uint8_t *q = p + i;
        for (int k = 0; k < rec; k++) q[k] = 0;
        if (rec >= 64) {
            uint32_t pos = serial * 3;
            memcpy(q + 8, &pos, 4);
            int l = snprintf((char *) q + 36, 28, "HWI-ST1234:100:C0:1:%u", 1000000 + serial);
            (void) l;
            for (int k = 64; k < rec && k < 64 + 75; k++) q[k] = (uint8_t) rnd();
            for (int k = 64 + 75; k < rec; k++) q[k] = (uint8_t) (30 + (rnd() % 11));
        }
        serial++;
        i += rec;
    }
}

int main(int argc, char **argv) {
    const size_t n_blocks = argc > 1 ? (size_t) atoll(argv[1]) : 4096;
    const int level = argc > 2 ? atoi(argv[2]) : 1;
    CK(cuInit(0));
    RT(cudaSetDevice(0));
    RT(cudaFree(0));
    CUdevice dev;
    CK(cuDeviceGet(&dev, 0));
    int mask = 0, maxlen = 0;
    CK(cuDeviceGetAttribute(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, dev));
    CK(cuDeviceGetAttribute(&maxlen, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH, dev));
    printf("{\"probe\": \"attributes\", \"algorithm_mask\": %d, \"max_length\": %d}\n", mask, maxlen);
    if (!(mask & CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE)) { printf("{\"probe\": \"unsupported\"}\n"); return 0; }

    // ---- input: n_blocks payloads of 65280 bytes (BGZF writers cut there or at 65536), raw deflate
    const int payload = 65280;
    std::vector<uint8_t> plain(n_blocks * (size_t) payload), comp;
    std::vector<uint64_t> in_off(n_blocks + 1), out_off(n_blocks + 1);
    std::vector<uint8_t> tmp(compressBound(payload) + 64);
    for (size_t b = 0; b < n_blocks; b++) {
        fill_block(&plain[b * payload], payload);
        z_stream z;
        memset(&z, 0, sizeof z);
        deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        z.next_in = &plain[b * payload];
        z.avail_in = payload;
        z.next_out = tmp.data();
        z.avail_out = (uInt) tmp.size();
        deflate(&z, Z_FINISH);
        const size_t cs = z.total_out;
        deflateEnd(&z);
        in_off[b] = comp.size() + 18;      // BGZF: 18 bytes of member header in front, 8 behind: odd alignment on purpose
        comp.insert(comp.end(), 18, 0);
        comp.insert(comp.end(), tmp.begin(), tmp.begin() + cs);
        comp.insert(comp.end(), 8, 0);
        out_off[b] = b * (size_t) payload;
    }
    in_off[n_blocks] = comp.size() + 18;
    out_off[n_blocks] = n_blocks * (size_t) payload;
    printf("{\"probe\": \"input\", \"blocks\": %zu, \"plain_bytes\": %zu, \"comp_bytes\": %zu, \"ratio\": %.3f}\n", n_blocks, plain.size(), comp.size(),
           (double) plain.size() / comp.size());

    uint8_t *d_comp, *d_out;
    uint32_t *d_act;
    const size_t lead = 37;      // destination misaligned too
    RT(cudaMalloc(&d_comp, comp.size() + 256));
    RT(cudaMalloc(&d_out, plain.size() + 256));
    RT(cudaMalloc(&d_act, n_blocks * 4));
    RT(cudaMemcpy(d_comp, comp.data(), comp.size(), cudaMemcpyHostToDevice));
    int capable = -1;
    cuPointerGetAttribute(&capable, CU_POINTER_ATTRIBUTE_IS_HW_DECOMPRESS_CAPABLE, (CUdeviceptr) d_out);
    printf("{\"probe\": \"pointer\", \"cudaMalloc_hw_decompress_capable\": %d}\n", capable);

    std::vector<CUmemDecompressParams> par(n_blocks);
    memset(par.data(), 0, par.size() * sizeof(CUmemDecompressParams));
    for (size_t b = 0; b < n_blocks; b++) {
        par[b].srcNumBytes = in_off[b + 1] - in_off[b] - 26;
        par[b].dstNumBytes = payload;
        par[b].dstActBytes = d_act + b;
        par[b].src = d_comp + in_off[b];
        par[b].dst = d_out + lead + out_off[b];
        par[b].algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
    }
    cudaStream_t s;
    RT(cudaStreamCreate(&s));
    cudaEvent_t e0, e1;
    RT(cudaEventCreate(&e0));
    RT(cudaEventCreate(&e1));
    std::vector<uint8_t> back(plain.size());
    std::vector<uint32_t> act(n_blocks);
    for (int rep = 0; rep < 4; rep++) {
        RT(cudaMemsetAsync(d_out, 0xEE, plain.size() + 256, s));
        RT(cudaMemsetAsync(d_act, 0, n_blocks * 4, s));
        RT(cudaEventRecord(e0, s));
        size_t err_index = (size_t) -1;
        // batches of at most `chunk` operations per call (rep 0: everything at once)
        const size_t chunk = rep < 2 ? n_blocks : (rep == 2 ? 8192 : 1024);
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (size_t i = 0; i < n_blocks; i += chunk) {
            const size_t m = n_blocks - i < chunk ? n_blocks - i : chunk;
            CUresult r = cuMemBatchDecompressAsync(par.data() + i, m, 0, &err_index, (CUstream) s);
            if (r != CUDA_SUCCESS) {
                const char *es = 0;
                cuGetErrorString(r, &es);
                printf("{\"error\": \"cuMemBatchDecompressAsync -> %d %s\", \"err_index\": %zu, \"rep\": %d}\n", (int) r, es ? es : "?", err_index, rep);
                return 2;
            }
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        RT(cudaEventRecord(e1, s));
        RT(cudaStreamSynchronize(s));
        float ms = 0;
        RT(cudaEventElapsedTime(&ms, e0, e1));
        RT(cudaMemcpy(back.data(), d_out + lead, plain.size(), cudaMemcpyDeviceToHost));
        RT(cudaMemcpy(act.data(), d_act, n_blocks * 4, cudaMemcpyDeviceToHost));
        size_t bad_bytes = 0, bad_act = 0;
        for (size_t b = 0; b < n_blocks; b++) {
            if (act[b] != (uint32_t) payload) bad_act++;
            if (memcmp(&back[b * payload], &plain[b * payload], payload)) bad_bytes++;
        }
        printf("{\"probe\": \"deflate\", \"rep\": %d, \"ops_per_call\": %zu, \"ms\": %.3f, \"submit_ms\": %.3f, \"GBps_out\": %.1f, \"GBps_in\": %.1f, \"blocks_wrong\": %zu, \"act_wrong\": %zu}\n",
               rep, chunk, ms, (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6, plain.size() / ms * 1e-6, comp.size() / ms * 1e-6, bad_bytes, bad_act);
    }
    const char *mode = argc > 3 ? argv[3] : "overlap";
    if (!strcmp(mode, "pipeline")) {
        // ---- the shape push_bgzf wants: upload piece k + 1 on one stream while the engine inflates piece k on another
        timespec a0, a1;
        uint8_t *h_pin;
        clock_gettime(CLOCK_MONOTONIC, &a0);
        RT(cudaMallocHost(&h_pin, comp.size()));
        clock_gettime(CLOCK_MONOTONIC, &a1);
        printf("{\"probe\": \"cudaMallocHost\", \"bytes\": %zu, \"ms\": %.2f}\n", comp.size(), (a1.tv_sec - a0.tv_sec) * 1e3 + (a1.tv_nsec - a0.tv_nsec) * 1e-6);
        memcpy(h_pin, comp.data(), comp.size());
        {
            void *big = 0;
            clock_gettime(CLOCK_MONOTONIC, &a0);
            RT(cudaMalloc(&big, (size_t) 10 << 30));
            clock_gettime(CLOCK_MONOTONIC, &a1);
            const double m = (a1.tv_sec - a0.tv_sec) * 1e3 + (a1.tv_nsec - a0.tv_nsec) * 1e-6;
            RT(cudaMemset(big, 1, (size_t) 10 << 30));
            RT(cudaDeviceSynchronize());
            clock_gettime(CLOCK_MONOTONIC, &a0);
            RT(cudaFree(big));
            clock_gettime(CLOCK_MONOTONIC, &a1);
            printf("{\"probe\": \"alloc\", \"cudaMalloc_10GB_ms\": %.2f, \"cudaFree_10GB_ms\": %.2f}\n", m, (a1.tv_sec - a0.tv_sec) * 1e3 + (a1.tv_nsec - a0.tv_nsec) * 1e-6);
        }
        cudaStream_t up;
        RT(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
        cudaStream_t s3;
        RT(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
        const size_t per = 1024;      // blocks per piece (~42 MB compressed)
        const size_t np = (n_blocks + per - 1) / per;
        std::vector<cudaEvent_t> up_done(np), inf_start(np), inf_done(np);
        for (size_t k = 0; k < np; k++) { RT(cudaEventCreate(&up_done[k])); RT(cudaEventCreate(&inf_start[k])); RT(cudaEventCreate(&inf_done[k])); }
        cudaEvent_t reused, z0;
        RT(cudaEventCreateWithFlags(&reused, cudaEventDisableTiming));
        RT(cudaEventCreate(&z0));
        for (int variant = 0; variant < 3; variant++) {      // 0: one event per piece, 1: one event re-recorded, 2: as 0 with batches of the whole piece split in 256-op calls
            RT(cudaMemset(d_comp, 0, comp.size()));      // stale data must not rescue a missing dependency
            RT(cudaMemsetAsync(d_act, 0, n_blocks * 4, s3));
            RT(cudaDeviceSynchronize());
            clock_gettime(CLOCK_MONOTONIC, &a0);
            RT(cudaEventRecord(z0, up));
            for (size_t k = 0; k < np; k++) {
                const size_t b0 = k * per, b1 = b0 + per < n_blocks ? b0 + per : n_blocks;
                const size_t lo = in_off[b0] - 18, hi = in_off[b1] - 18;
                RT(cudaMemcpyAsync(d_comp + lo, h_pin + lo, hi - lo, cudaMemcpyHostToDevice, up));
                RT(cudaEventRecord(up_done[k], up));
                if (variant == 1) { RT(cudaEventRecord(reused, up)); RT(cudaStreamWaitEvent(s3, reused, 0)); }
                else RT(cudaStreamWaitEvent(s3, up_done[k], 0));
                RT(cudaEventRecord(inf_start[k], s3));
                size_t err_index = (size_t) -1;
                const size_t call = variant == 2 ? 256 : per;
                for (size_t i = b0; i < b1; i += call) CK(cuMemBatchDecompressAsync(par.data() + i, b1 - i < call ? b1 - i : call, 0, &err_index, (CUstream) s3));
                RT(cudaEventRecord(inf_done[k], s3));
            }
            clock_gettime(CLOCK_MONOTONIC, &a1);
            RT(cudaDeviceSynchronize());
            RT(cudaMemcpy(act.data(), d_act, n_blocks * 4, cudaMemcpyDeviceToHost));
            size_t bad_act = 0;
            for (size_t b = 0; b < n_blocks; b++) bad_act += act[b] != (uint32_t) payload;
            float t_up_last = 0, t_inf_first = 0, t_inf_last = 0, t_mid_up = 0, t_mid_inf0 = 0, t_mid_inf1 = 0;
            RT(cudaEventElapsedTime(&t_up_last, z0, up_done[np - 1]));
            RT(cudaEventElapsedTime(&t_inf_first, z0, inf_start[0]));
            RT(cudaEventElapsedTime(&t_inf_last, z0, inf_done[np - 1]));
            RT(cudaEventElapsedTime(&t_mid_up, z0, up_done[np / 2]));
            RT(cudaEventElapsedTime(&t_mid_inf0, z0, inf_start[np / 2]));
            RT(cudaEventElapsedTime(&t_mid_inf1, z0, inf_done[np / 2]));
            printf("{\"probe\": \"pipeline\", \"variant\": %d, \"pieces\": %zu, \"host_issue_ms\": %.2f, \"upload_done_ms\": %.2f, \"first_inflate_start_ms\": %.2f, \"all_done_ms\": %.2f, "
                   "\"mid_piece\": [%.2f, %.2f, %.2f], \"act_wrong\": %zu, \"h2d_GBps\": %.1f}\n", variant, np, (a1.tv_sec - a0.tv_sec) * 1e3 + (a1.tv_nsec - a0.tv_nsec) * 1e-6,
                   t_up_last, t_inf_first, t_inf_last, t_mid_up, t_mid_inf0, t_mid_inf1, bad_act, comp.size() / t_up_last * 1e-6);
        }
    } else
    if (!strcmp(mode, "overlap")) {
        // ---- does the engine run next to an H2D copy on another stream (the chunked upload of the compressed file)?
        const size_t hb = (size_t) 2 << 30;
        uint8_t *h_pin, *d_side;
        RT(cudaMallocHost(&h_pin, hb));
        memset(h_pin, 1, hb);
        RT(cudaMalloc(&d_side, hb));
        cudaStream_t s2;
        RT(cudaStreamCreate(&s2));
        cudaEvent_t c0, c1;
        RT(cudaEventCreate(&c0));
        RT(cudaEventCreate(&c1));
        for (int both = 0; both < 2; both++) {
            RT(cudaDeviceSynchronize());
            RT(cudaEventRecord(c0, s2));
            RT(cudaMemcpyAsync(d_side, h_pin, hb, cudaMemcpyHostToDevice, s2));
            RT(cudaEventRecord(c1, s2));
            RT(cudaEventRecord(e0, s));
            if (both) {
                size_t err_index = (size_t) -1;
                for (int k = 0; k < 2; k++) CK(cuMemBatchDecompressAsync(par.data(), n_blocks, 0, &err_index, (CUstream) s));
            }
            RT(cudaEventRecord(e1, s));
            RT(cudaDeviceSynchronize());
            float mc = 0, md = 0;
            RT(cudaEventElapsedTime(&mc, c0, c1));
            RT(cudaEventElapsedTime(&md, e0, e1));
            printf("{\"probe\": \"overlap\", \"with_engine\": %d, \"h2d_ms\": %.3f, \"h2d_GBps\": %.1f, \"engine_ms\": %.3f, \"engine_GBps_out\": %.1f}\n", both, mc, hb / mc * 1e-6, md,
                   both ? 2.0 * plain.size() / md * 1e-6 : 0.0);
        }
    } else if (!strcmp(mode, "short_dst")) {
        // ---- a valid stream whose dstNumBytes understates the output: is the destination bounded by it?
        RT(cudaMemsetAsync(d_out, 0xEE, plain.size() + 256, s));
        RT(cudaMemsetAsync(d_act, 0xFF, 8, s));
        par[0].dstNumBytes = 1000;
        size_t err_index = (size_t) -1;
        CUresult r = cuMemBatchDecompressAsync(par.data(), 1, 0, &err_index, (CUstream) s);
        cudaError_t se = cudaStreamSynchronize(s);
        uint32_t a = 0;
        std::vector<uint8_t> b2(payload);
        cudaMemcpy(&a, d_act, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(b2.data(), d_out + lead, payload, cudaMemcpyDeviceToHost);
        size_t written = 0;
        for (int k = 0; k < payload; k++) if (b2[k] == plain[k] && b2[k] != 0xEE) written = k + 1;
        void *again = 0;
        cudaError_t ma = cudaMalloc(&again, 1 << 20);
        printf("{\"probe\": \"short_dst\", \"submit\": %d, \"sync\": \"%s\", \"act_bytes\": %u, \"last_matching_byte\": %zu, \"context_alive\": \"%s\"}\n", (int) r, cudaGetErrorString(se), a, written,
               cudaGetErrorString(ma));
    } else {
        // ---- what a corrupt stream does (the library must report "Zlib inflate failed" like the reference)
        std::vector<uint8_t> bad(comp.begin(), comp.begin() + (size_t) in_off[1]);
        if (!strcmp(mode, "truncated")) par[0].srcNumBytes /= 2;
        else for (size_t k = in_off[0] + 40; k < in_off[0] + 60 && k < bad.size(); k++) bad[k] ^= 0xA5;
        RT(cudaMemcpy(d_comp, bad.data(), bad.size(), cudaMemcpyHostToDevice));
        RT(cudaMemsetAsync(d_act, 0xFF, 4, s));
        size_t err_index = (size_t) -1;
        CUresult r = cuMemBatchDecompressAsync(par.data(), 1, 0, &err_index, (CUstream) s);
        cudaError_t se = cudaStreamSynchronize(s);
        uint32_t a = 0;
        cudaError_t ce = cudaMemcpy(&a, d_act, 4, cudaMemcpyDeviceToHost);
        void *again = 0;
        cudaError_t ma = cudaMalloc(&again, 1 << 20);
        printf("{\"probe\": \"%s\", \"submit\": %d, \"sync\": \"%s\", \"act_bytes\": %u, \"copy_after\": \"%s\", \"malloc_after\": \"%s\"}\n", mode, (int) r, cudaGetErrorString(se), a,
               cudaGetErrorString(ce), cudaGetErrorString(ma));
    }
    return 0;
}
