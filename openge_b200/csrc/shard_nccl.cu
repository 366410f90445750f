// The range-sharded step driven from C++: the five phases of shard_api.cu with the four exchanges between them done by
// NCCL directly -- grouped ncclSend / ncclRecv per peer, i.e. an all-to-all with uneven splits, on the library's own
// stream.  Nothing crosses into a host language between the phases; per exchange the host waits once, for the item counts
// that size the receive buffers.  The lists leave straight from the buffers the bucketing wrote (no packing); the key hashes
// that go to every rank are sent from the one copy.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy the process already has -- torch's under torch.distributed, the
// system's in a plain C++ program), so the library neither links it nor needs it on one GPU.
#include <dlfcn.h>
#include <string.h>

#include <vector>

#include "ctx.cuh"

namespace {

// the part of nccl.h this file uses (NCCL keeps these signatures stable across 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1 };

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.lib) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail_msg(OGE_ERR_CUDA, "NCCL is not available: %s", dlerror());
    NcclApi a;
    a.lib = h;
    *(void **) &a.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void **) &a.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void **) &a.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void **) &a.GroupStart = dlsym(h, "ncclGroupStart");
    *(void **) &a.GroupEnd = dlsym(h, "ncclGroupEnd");
    *(void **) &a.Send = dlsym(h, "ncclSend");
    *(void **) &a.Recv = dlsym(h, "ncclRecv");
    *(void **) &a.GetErrorString = dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.GroupStart || !a.GroupEnd || !a.Send || !a.Recv || !a.GetErrorString)
        return fail_msg(OGE_ERR_CUDA, "NCCL: a symbol is missing from libnccl.so.2");
    g_nccl = a;
    return 0;
}

#define OGE_NCCL_TRY(expr)                                                                                                   \
    do {                                                                                                                     \
        ncclResult_t _r = (expr);                                                                                            \
        if (_r != 0) return fail_msg(OGE_ERR_CUDA, "NCCL error %d (%s) at %s:%d: %s", _r, g_nccl.GetErrorString(_r), __FILE__, __LINE__, #expr); \
    } while (0)

// One list of an exchange.  Sending side: `send` ordered by destination with send_counts[d] items for rank d -- or, with
// `to_all`, the same n_all items for every OTHER rank.  Receiving side: everything addressed to this rank, in rank order.
struct XList {
    const void *send = nullptr;
    const uint64_t *send_counts = nullptr;
    bool to_all = false;
    uint64_t n_all = 0;
    uint32_t item = 0;
    DevBuf<uint8_t> *recv = nullptr;
    uint64_t recv_items = 0;
};

int nccl_exchange(oge_gpu_dedup_ctx *c, XList *L, int k) {
    ShardState &sh = c->sh;
    cudaStream_t s = c->stream;
    const int W = c->cfg.world, me = c->cfg.rank;
    ncclComm_t comm = (ncclComm_t) sh.nccl_comm;
    int rc;
    // ---- the item counts first: cnt[d * k + j] = items of list j this rank has for rank d
    if ((rc = sh.x_cnt.reserve(2 * (size_t) W * k, false, s))) return rc;
    std::vector<uint64_t> h_send((size_t) W * k), h_recv((size_t) W * k);
    for (int d = 0; d < W; d++)
        for (int j = 0; j < k; j++) h_send[(size_t) d * k + j] = L[j].to_all ? (d == me ? 0 : L[j].n_all) : L[j].send_counts[d];
    uint64_t *d_send = sh.x_cnt.p, *d_recv = sh.x_cnt.p + (size_t) W * k;
    OGE_CUDA_TRY(cudaMemcpyAsync(d_send, h_send.data(), h_send.size() * 8, cudaMemcpyHostToDevice, s));
    OGE_NCCL_TRY(g_nccl.GroupStart());
    for (int p = 0; p < W; p++) {
        OGE_NCCL_TRY(g_nccl.Send(d_send + (size_t) p * k, (size_t) k * 8, ncclUint8, p, comm, s));
        OGE_NCCL_TRY(g_nccl.Recv(d_recv + (size_t) p * k, (size_t) k * 8, ncclUint8, p, comm, s));
    }
    OGE_NCCL_TRY(g_nccl.GroupEnd());
    OGE_CUDA_TRY(cudaMemcpyAsync(h_recv.data(), d_recv, h_recv.size() * 8, cudaMemcpyDeviceToHost, s));
    OGE_CUDA_TRY(cudaStreamSynchronize(s));
    // ---- the payloads, straight out of the bucketed buffers
    for (int j = 0; j < k; j++) {
        uint64_t tot = 0;
        for (int p = 0; p < W; p++) tot += h_recv[(size_t) p * k + j];
        L[j].recv_items = tot;
        if ((rc = L[j].recv->reserve(tot * L[j].item + 16, false, s))) return rc;
    }
    OGE_NCCL_TRY(g_nccl.GroupStart());
    for (int j = 0; j < k; j++) {
        uint64_t soff = 0, roff = 0;
        for (int p = 0; p < W; p++) {
            const uint64_t ns = h_send[(size_t) p * k + j], nr = h_recv[(size_t) p * k + j];
            if (ns) OGE_NCCL_TRY(g_nccl.Send((const uint8_t *) L[j].send + (L[j].to_all ? 0 : soff * L[j].item), ns * L[j].item, ncclUint8, p, comm, s));
            if (nr) OGE_NCCL_TRY(g_nccl.Recv(L[j].recv->p + roff * L[j].item, nr * L[j].item, ncclUint8, p, comm, s));
            soff += ns;
            roff += nr;
            sh.x_bytes += p == me ? 0 : ns * L[j].item;
        }
    }
    OGE_NCCL_TRY(g_nccl.GroupEnd());
    sh.x_calls++;
    return 0;
}

}  // namespace

extern "C" {

int oge_gpu_shard_comm_id(uint8_t *id128) {
    if (!id128) return fail_msg(OGE_ERR_INVALID_ARG, "shard_comm_id: null argument");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    OGE_NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return OGE_OK;
}

int oge_gpu_shard_comm_init(oge_gpu_dedup_ctx *c, const uint8_t *id128) {
    if (!c || !id128) return fail_msg(OGE_ERR_INVALID_ARG, "shard_comm_init: null argument");
    if (c->cfg.world < 1 || c->cfg.rank < 0 || c->cfg.rank >= c->cfg.world) return fail_msg(OGE_ERR_INVALID_ARG, "shard_comm_init: rank %d of %d", c->cfg.rank, c->cfg.world);
    int rc = nccl_load();
    if (rc) return rc;
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    if (c->sh.nccl_comm) {
        g_nccl.CommDestroy((ncclComm_t) c->sh.nccl_comm);
        c->sh.nccl_comm = nullptr;
    }
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclComm_t comm = nullptr;
    OGE_NCCL_TRY(g_nccl.CommInitRank(&comm, c->cfg.world, id, c->cfg.rank));
    c->sh.nccl_comm = comm;
    return OGE_OK;
}

void oge_gpu_shard_comm_destroy(oge_gpu_dedup_ctx *c) {
    if (c && c->sh.nccl_comm && g_nccl.lib) {
        cudaSetDevice(c->cfg.device);
        g_nccl.CommDestroy((ncclComm_t) c->sh.nccl_comm);
        c->sh.nccl_comm = nullptr;
    }
}

int oge_gpu_shard_step(oge_gpu_dedup_ctx *c, oge_gpu_shard_step_info *info) {
    if (!c) return fail_msg(OGE_ERR_INVALID_ARG, "shard_step: null context");
    if (!c->sh.nccl_comm) return fail_msg(OGE_ERR_STATE, "shard_step: call oge_gpu_shard_comm_init first");
    ShardState &sh = c->sh;
    const int W = c->cfg.world;
    std::vector<uint64_t> ca(W), cb(W);
    void *pa = nullptr, *pb = nullptr, *ph = nullptr;
    uint64_t nh = 0;
    int rc;
    sh.x_bytes = 0;
    sh.x_calls = 0;
    // begin -> published entries to the name owners, their hashes to all, boundary fragment ends to the key owners
    if ((rc = oge_gpu_shard_begin(c, &pa, ca.data(), &ph, &nh, &pb, cb.data()))) return rc;
    XList x1[3];
    x1[0].send = pa; x1[0].send_counts = ca.data(); x1[0].item = sh.entry_bytes; x1[0].recv = &sh.r_pub;
    x1[1].send = pb; x1[1].send_counts = cb.data(); x1[1].item = sizeof(RouteEntry); x1[1].recv = &sh.r_froute;
    x1[2].send = ph; x1[2].to_all = true; x1[2].n_all = nh; x1[2].item = 8; x1[2].recv = &sh.r_hash;
    if ((rc = nccl_exchange(c, x1, 3))) return rc;
    const uint64_t n_pub1 = x1[0].recv_items;
    // probe -> round-2 entries to the name owners, local pair ends to the key owners
    if ((rc = oge_gpu_shard_probe(c, sh.r_hash.p, x1[2].recv_items, sh.r_froute.p, x1[1].recv_items, &pa, ca.data(), &pb, cb.data()))) return rc;
    XList x2[2];
    x2[0].send = pa; x2[0].send_counts = ca.data(); x2[0].item = sh.entry_bytes; x2[0].recv = &sh.r_pub2;
    x2[1].send = pb; x2[1].send_counts = cb.data(); x2[1].item = sizeof(RouteEntry); x2[1].recv = &sh.r_proute;
    if ((rc = nccl_exchange(c, x2, 2))) return rc;
    const uint64_t n_pub2 = x2[0].recv_items, n_pr = x2[1].recv_items;
    // replay over both rounds (round 2 appended behind round 1)
    if (n_pub2) {
        if ((rc = sh.r_pub.reserve((n_pub1 + n_pub2) * sh.entry_bytes + 16, true, c->stream))) return rc;
        OGE_CUDA_TRY(cudaMemcpyAsync(sh.r_pub.p + n_pub1 * sh.entry_bytes, sh.r_pub2.p, n_pub2 * sh.entry_bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    if ((rc = oge_gpu_shard_replay(c, sh.r_pub.p, n_pub1 + n_pub2, &pa, ca.data()))) return rc;
    XList x3[1];
    x3[0].send = pa; x3[0].send_counts = ca.data(); x3[0].item = sizeof(RouteEntry); x3[0].recv = &sh.r_oroute;
    if ((rc = nccl_exchange(c, x3, 1))) return rc;
    const uint64_t n_or = x3[0].recv_items;
    // finish over the pair ends of both kinds (the replayed ones appended behind the local ones)
    if (n_or) {
        if ((rc = sh.r_proute.reserve((n_pr + n_or) * sizeof(RouteEntry) + 16, true, c->stream))) return rc;
        OGE_CUDA_TRY(cudaMemcpyAsync(sh.r_proute.p + n_pr * sizeof(RouteEntry), sh.r_oroute.p, n_or * sizeof(RouteEntry), cudaMemcpyDeviceToDevice, c->stream));
    }
    if ((rc = oge_gpu_shard_finish(c, sh.r_proute.p, n_pr + n_or, &pa, ca.data()))) return rc;
    XList x4[1];
    x4[0].send = pa; x4[0].send_counts = ca.data(); x4[0].item = 4; x4[0].recv = &sh.r_marks;
    if ((rc = nccl_exchange(c, x4, 1))) return rc;
    if ((rc = oge_gpu_shard_apply(c, sh.r_marks.p, x4[0].recv_items))) return rc;
    if (info) {
        info->published_in = n_pub1 + n_pub2;
        info->routed_in = x1[1].recv_items + n_pr + n_or;
        info->marks_in = x4[0].recv_items;
        info->exchanges = sh.x_calls;
        info->bytes_sent = sh.x_bytes;
    }
    return OGE_OK;
}

}  // extern "C"
