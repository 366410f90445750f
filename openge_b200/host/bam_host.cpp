// Host-side BAM streaming layer around the GPU dedup path: include/oge_bam_host.h (SURVEY 8(f) f1 + f2).
//
// The reference moves one heap-allocated OGERead per record through FileReader -> MarkDuplicates -> FileWriter
// (commands/command_dedup.cpp:48-69), inflating BGZF blocks through a job queue polled every 50 ms
// (util/bgzf_input_stream.cpp:217) and deserialising on one thread.  Here a file is three flat things: the inflated
// stream in one (pinned) buffer, an offsets array, a parsed header.  BGZF blocks are independent, so both
// directions run as plain parallel loops over blocks with no queue.
//
// Byte-exactness: the store side cuts the stream and calls zlib exactly as util/bgzf_output_stream.cpp does
// (65536-byte blocks, 65472 at level 0; deflateInit2(level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) into
// 65536 - 26 bytes, level + 1 on overflow; the same 18 header bytes; a final partial block -- even an empty one --
// followed by an empty block), so with the same zlib the output FILE equals the reference's, not only its
// decompressed content.
#include "oge_bam_host.h"

#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/mman.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

inline uint16_t rd_u16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rd_i32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
inline void wr_u16(uint8_t *p, uint16_t v) { memcpy(p, &v, 2); }
inline void wr_u32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }

int clamp_threads(int t) {
    if (t <= 0) t = (int) std::thread::hardware_concurrency();
    return std::max(1, std::min(t, 256));
}

// fn(worker) on `threads` workers; the first non-zero return code wins (its message is copied to the caller's g_err)
template <typename F>
int parallel_run(int threads, F fn) {
    std::vector<std::thread> pool;
    std::vector<int> rcs(threads, 0);
    std::vector<std::string> msgs(threads);
    for (int w = 0; w < threads; w++)
        pool.emplace_back([&, w] {
            rcs[w] = fn(w);
            if (rcs[w]) msgs[w] = g_err;
        });
    for (auto &t : pool) t.join();
    for (int w = 0; w < threads; w++)
        if (rcs[w]) return fail(rcs[w], "%s", msgs[w].c_str());
    return 0;
}

// ------------------------------------------------------------------------------------------------ BGZF, read side
constexpr uint32_t BGZF_BLOCK = 65536;      // util/bgzf_output_stream.h:26, util/bgzf_input_stream.cpp:113

struct BlockIndex {
    std::vector<uint64_t> in_off, out_off;      // out_off has one more entry: the total
    std::vector<uint32_t> csize, isize;
};

// Header checks of BgzfInputStream::BgzfBlock::decompress (util/bgzf_input_stream.cpp:76-98).
int bgzf_scan(const uint8_t *d, size_t n, BlockIndex &ix) {
    size_t pos = 0;
    uint64_t total = 0;
    while (pos < n) {
        if (n - pos < 18) return fail(OGE_BAM_ERR_FORMAT, "BGZF: truncated block header at byte %zu", pos);
        const uint8_t *h = d + pos;
        if (h[0] != 31 || h[1] != 139) return fail(OGE_BAM_ERR_FORMAT, "BGZF block has invalid start block. Is this file corrupted?");
        if (h[2] != 8 || h[3] != 4) return fail(OGE_BAM_ERR_FORMAT, "BGZF block has unexpected flags. Is this file corrupted?");
        if (rd_u16(h + 10) != 6 || h[12] != 66 || h[13] != 67) return fail(OGE_BAM_ERR_FORMAT, "BGZF GZ extra field is incorrect. Is this file corrupted?");
        const uint32_t bsize = (uint32_t) rd_u16(h + 16) + 1;
        if (bsize < 26 || pos + bsize > n) return fail(OGE_BAM_ERR_FORMAT, "BGZF: block of %u bytes at byte %zu overruns the file", bsize, pos);
        const uint32_t isize = rd_u32(h + bsize - 4);
        if (isize > BGZF_BLOCK) return fail(OGE_BAM_ERR_FORMAT, "BGZF: block inflates to %u bytes (more than 65536)", isize);
        ix.in_off.push_back(pos);
        ix.csize.push_back(bsize);
        ix.isize.push_back(isize);
        ix.out_off.push_back(total);
        total += isize;
        pos += bsize;
    }
    ix.out_off.push_back(total);
    return 0;
}

int bgzf_inflate(const uint8_t *d, const BlockIndex &ix, uint8_t *out, int threads) {
    const size_t nb = ix.in_off.size();
    std::atomic<size_t> next(0);
    threads = (int) std::min<size_t>(threads, std::max<size_t>(1, nb / 4));
    return parallel_run(threads, [&](int) -> int {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) return fail(OGE_BAM_ERR_NOMEM, "Zlib initialization failed.");
        int rc = 0;
        while (!rc) {
            const size_t b0 = next.fetch_add(16);
            if (b0 >= nb) break;
            for (size_t b = b0; b < std::min(nb, b0 + 16) && !rc; b++) {
                if (ix.isize[b] == 0) continue;
                inflateReset(&zs);
                zs.next_in = const_cast<Bytef *>(d + ix.in_off[b] + 18);
                zs.avail_in = ix.csize[b] - 18;      // the reference hands zlib the footer as well (:110); it stops at the end of the stream
                zs.next_out = out + ix.out_off[b];
                zs.avail_out = ix.isize[b];
                const int st = inflate(&zs, Z_FINISH);
                if (st != Z_STREAM_END || zs.total_out != ix.isize[b]) rc = fail(OGE_BAM_ERR_FORMAT, "Zlib inflate failed (BGZF block %zu).", b);
            }
        }
        inflateEnd(&zs);
        return rc;
    });
}

// ------------------------------------------------------------------------------------------------ BGZF, write side
// One block exactly as BgzfOutputStream::BgzfBlock::compress (util/bgzf_output_stream.cpp:59-141).  dst: 65536 bytes.
int bgzf_compress_block(const uint8_t *src, uint32_t len, int level, uint8_t *dst, uint32_t *out_len) {
    int cur = level;
    uLong produced = 0;
    while (true) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        zs.next_in = const_cast<Bytef *>(src);
        zs.avail_in = len;
        zs.next_out = dst + 18;
        zs.avail_out = BGZF_BLOCK - 18 - 8;
        if (deflateInit2(&zs, cur, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return fail(OGE_BAM_ERR_NOMEM, "BGZF writer: zlib deflateInit2 failed");
        const int st = deflate(&zs, Z_FINISH);
        produced = zs.total_out;
        if (deflateEnd(&zs) != Z_OK && st == Z_STREAM_END) return fail(OGE_BAM_ERR_FORMAT, "BGZF writer: zlib deflateEnd failed");
        if (st == Z_STREAM_END) break;
        if (st == Z_OK) {      // did not fit: the reference retries one compression level up (:103-111)
            if (++cur > Z_BEST_COMPRESSION) return fail(OGE_BAM_ERR_FORMAT, "BGZF writer: input reduction failed");
            continue;
        }
        return fail(OGE_BAM_ERR_FORMAT, "BGZF writer: zlib deflate failed");
    }
    const uint32_t csize = (uint32_t) produced + 18 + 8;
    static const uint8_t head[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0};      // :117-131
    memcpy(dst, head, 16);
    wr_u16(dst + 16, (uint16_t) (csize - 1));
    wr_u32(dst + csize - 8, (uint32_t) crc32(crc32(0, NULL, 0), src, len));
    wr_u32(dst + csize - 4, len);
    *out_len = csize;
    return 0;
}

// A byte stream made of segments (header bytes, then the record buffer), read without gluing it together.
struct Segments {
    std::vector<const uint8_t *> ptr;
    std::vector<uint64_t> start;      // stream offset of each segment; one more entry: the total
    void add(const uint8_t *p, uint64_t n) {
        if (start.empty()) start.push_back(0);
        ptr.push_back(p);
        start.push_back(start.back() + n);
    }
    uint64_t total() const { return start.empty() ? 0 : start.back(); }
    // -> pointer to `len` contiguous bytes at stream offset `at` (copied into tmp only when they straddle segments)
    const uint8_t *span(uint64_t at, uint32_t len, uint8_t *tmp) const {
        size_t s = std::upper_bound(start.begin(), start.end(), at) - start.begin() - 1;
        if (s >= ptr.size()) return tmp;      // len == 0 at the very end
        if (at + len <= start[s + 1]) return ptr[s] + (at - start[s]);
        uint32_t done = 0;
        while (done < len) {
            const uint64_t avail = start[s + 1] - (at + done);
            const uint32_t take = (uint32_t) std::min<uint64_t>(avail, len - done);
            memcpy(tmp + done, ptr[s] + (at + done - start[s]), take);
            done += take;
            s++;
        }
        return tmp;
    }
};

struct Sink {      // where compressed bytes go: a file descriptor or a growing buffer
    int fd = -1;
    std::vector<uint8_t> *buf = nullptr;
    int write(const uint8_t *p, size_t n) {
        if (buf) {
            buf->insert(buf->end(), p, p + n);
            return 0;
        }
        while (n) {
            const ssize_t w = ::write(fd, p, n);
            if (w < 0) return fail(OGE_BAM_ERR_IO, "write failed: %s", strerror(errno));
            p += w;
            n -= (size_t) w;
        }
        return 0;
    }
};

// The block sequence of BgzfOutputStream::write + close (:170-250): full blocks, the current block (whatever it
// holds, possibly nothing), an empty block.  Blocks are compressed in waves of WAVE blocks by all threads, and a
// wave is written out while the next one is being compressed.
int bgzf_deflate_stream(const Segments &in, int level, int threads, Sink &sink) {
    const uint32_t full = level == 0 ? BGZF_BLOCK - 64 : BGZF_BLOCK;      // :143-146
    const uint64_t total = in.total();
    const uint64_t n_data_blocks = total / full + 1;      // the last one is the partial (maybe empty) block
    const uint64_t n_blocks = n_data_blocks + 1;          // + the empty block of close()
    const size_t WAVE = 1024;
    std::vector<uint8_t> arena[2];
    std::vector<uint32_t> sizes[2];
    arena[0].resize(std::min<uint64_t>(WAVE, n_blocks) * BGZF_BLOCK);
    arena[1].resize(n_blocks > WAVE ? WAVE * BGZF_BLOCK : 0);
    sizes[0].resize(WAVE);
    sizes[1].resize(WAVE);
    std::thread writer;
    int writer_rc = 0;
    std::string writer_msg;
    int rc = 0;
    int cur = 0;
    for (uint64_t w0 = 0; w0 < n_blocks && !rc; w0 += WAVE, cur ^= 1) {
        const size_t nw = (size_t) std::min<uint64_t>(WAVE, n_blocks - w0);
        std::atomic<size_t> next(0);
        uint8_t *ar = arena[cur].data();
        uint32_t *sz = sizes[cur].data();
        rc = parallel_run((int) std::min<size_t>(threads, nw), [&](int) -> int {
            std::vector<uint8_t> tmp(BGZF_BLOCK);
            while (true) {
                const size_t k = next.fetch_add(1);
                if (k >= nw) return 0;
                const uint64_t b = w0 + k;
                const uint64_t at = std::min(b * (uint64_t) full, total);
                const uint32_t len = b < n_data_blocks ? (uint32_t) std::min<uint64_t>(full, total - at) : 0;
                const uint8_t *src = in.span(at, len, tmp.data());
                const int r = bgzf_compress_block(src, len, level, ar + k * BGZF_BLOCK, &sz[k]);
                if (r) return r;
            }
        });
        if (writer.joinable()) writer.join();      // the previous wave is on disk (or failed)
        if (!rc && writer_rc) rc = fail(writer_rc, "%s", writer_msg.c_str());
        if (rc) break;
        writer = std::thread([&sink, ar, sz, nw, &writer_rc, &writer_msg] {
            for (size_t k = 0; k < nw && !writer_rc; k++) {
                writer_rc = sink.write(ar + k * BGZF_BLOCK, sz[k]);
                if (writer_rc) writer_msg = g_err;
            }
        });
    }
    if (writer.joinable()) writer.join();
    if (!rc && writer_rc) rc = fail(writer_rc, "%s", writer_msg.c_str());
    return rc;
}

// ------------------------------------------------------------------------------------------------ header model
// BamHeader(text) and toString() of the reference (util/bam_header.cpp:24-262, util/bam_header.h), field by field.
std::vector<std::string> header_line_split(const std::string &line) {      // :27-40
    std::vector<std::string> ret;
    size_t i = 0;
    while (true) {
        const size_t t = line.find('\t', i);
        if (t == std::string::npos) {
            ret.push_back(line.substr(i));
            break;
        }
        ret.push_back(line.substr(i, t - i));
        i = t + 1;
    }
    return ret;
}

struct SqRec {
    std::string name, as, m5, sp, ur;
    long long length = -1;
};
struct RgRec {
    std::string id, cn, ds, dt, fo, ks, lb, pg, pi, pl, pu, sm;
};
struct PgRec {
    std::string id, pn, cl, pp, vn;
};

struct HeaderModel {
    std::string version;
    int sort = 3;      // 0 unknown, 1 unsorted, 2 queryname, 3 coordinate (only the rendering matters)
    std::vector<SqRec> sq;
    std::vector<RgRec> rg;
    std::vector<PgRec> pg;
    std::vector<std::string> co;

    int parse(const std::string &text) {
        size_t i = 0;
        bool have_hd = false;
        while (true) {
            // getline + "if(!in.good()) break" (:111-116): a last line without '\n' is DROPPED
            const size_t nl = text.find('\n', i);
            if (nl == std::string::npos) break;
            const std::string line = text.substr(i, nl - i);
            i = nl + 1;
            if (line.empty() || line[0] != '@') return fail(OGE_BAM_ERR_FORMAT, "Sam header format problem: line doesn't begin with a '@'.");
            if (line.size() < 4 || line[3] != '\t') return fail(OGE_BAM_ERR_FORMAT, "Sam header format problem: line doesn't have a tab after the tag.");
            const std::string tag = line.substr(1, 2), data = line.substr(4);
            if (tag == "CO") {
                co.push_back(data);
                continue;
            }
            const std::vector<std::string> segs = header_line_split(data);
            for (const std::string &s : segs)
                if (s.size() < 3) return fail(OGE_BAM_ERR_FORMAT, "Sam header format problem: field '%s' is too short.", s.c_str());
            if (tag == "RG") {
                RgRec r;
                for (const std::string &s : segs) {
                    const std::string t = s.substr(0, 2), d = s.substr(3);
                    if (t == "ID") r.id = d; else if (t == "CN") r.cn = d; else if (t == "DS") r.ds = d; else if (t == "DT") r.dt = d;
                    else if (t == "FO") r.fo = d; else if (t == "KS") r.ks = d; else if (t == "LB") r.lb = d; else if (t == "PG") r.pg = d;
                    else if (t == "PI") r.pi = d; else if (t == "PL") r.pl = d; else if (t == "PU") r.pu = d; else if (t == "SM") r.sm = d;
                }
                if (r.id.empty()) return fail(OGE_BAM_ERR_FORMAT, "Mandatory field missing in header read group line.");
                rg.push_back(r);
            } else if (tag == "SQ") {
                SqRec r;
                for (const std::string &s : segs) {
                    const std::string t = s.substr(0, 2), d = s.substr(3);
                    if (t == "SN") r.name = d; else if (t == "LN") r.length = atoi(d.c_str()); else if (t == "AS") r.as = d;
                    else if (t == "M5") r.m5 = d; else if (t == "SP") r.sp = d; else if (t == "UR") r.ur = d;
                }
                if (r.name.empty() || r.length == -1) return fail(OGE_BAM_ERR_FORMAT, "Mandatory field missing in header sequence line.");
                sq.push_back(r);
            } else if (tag == "PG") {
                PgRec r;
                for (const std::string &s : segs) {
                    const std::string t = s.substr(0, 2), d = s.substr(3);
                    if (t == "ID") r.id = d; else if (t == "PN") r.pn = d; else if (t == "CL") r.cl = d; else if (t == "PP") r.pp = d;
                    else if (t == "VN") r.vn = d;
                }
                if (r.id.empty()) return fail(OGE_BAM_ERR_FORMAT, "Mandatory field missing in header program record line.");
                pg.push_back(r);
            } else if (tag == "HD") {
                std::string so, vn;
                for (const std::string &s : segs) {
                    const std::string t = s.substr(0, 2), d = s.substr(3);
                    if (t == "VN") vn = d; else if (t == "SO") so = d;
                }
                if (so.empty() || vn.empty()) return fail(OGE_BAM_ERR_FORMAT, "Mandatory field missing in header HD line.");
                version = vn;
                if (so == "unsorted") sort = 1; else if (so == "coordinate") sort = 3; else if (so == "queryname") sort = 2;
                else if (so == "unknown") sort = 0; else return fail(OGE_BAM_ERR_FORMAT, "Unknown sort order '%s'.", so.c_str());
                have_hd = true;
            } else {
                return fail(OGE_BAM_ERR_FORMAT, "Sam header format problem: tag '%s' wasn't CO RG SQ PG or HD.", tag.c_str());
            }
        }
        if (!have_hd && version.empty()) {      // :176-179
            version = "1.4";
            sort = 0;
        }
        return 0;
    }

    std::string render() const {      // :184-262
        static const char *so[] = {"unknown", "unsorted", "queryname", "coordinate"};
        std::string s = "@HD\tVN:" + version + "\tSO:" + so[sort] + "\n";
        for (const SqRec &r : sq) {
            s += "@SQ\tSN:" + r.name + "\tLN:" + std::to_string((unsigned long long) (size_t) r.length);
            if (!r.as.empty()) s += "\tAS:" + r.as;
            if (!r.m5.empty()) s += "\tM5:" + r.m5;
            if (!r.sp.empty()) s += "\tSP:" + r.sp;
            if (!r.ur.empty()) s += "\tUR:" + r.ur;
            s += "\n";
        }
        for (const RgRec &r : rg) {
            s += "@RG\tID:" + r.id;
            if (!r.cn.empty()) s += "\tCN:" + r.cn;
            if (!r.ds.empty()) s += "\tDS:" + r.ds;
            if (!r.dt.empty()) s += "\tDT:" + r.dt;
            if (!r.fo.empty()) s += "\tFO:" + r.fo;
            if (!r.ks.empty()) s += "\tKS:" + r.ks + "\tKS:" + r.ks;      // printed twice by the reference (:243-246)
            if (!r.lb.empty()) s += "\tLB:" + r.lb;
            if (!r.pg.empty()) s += "\tPG:" + r.pg;
            if (!r.pi.empty()) s += "\tPI:" + r.pi;
            if (!r.pl.empty()) s += "\tPL:" + r.pl;
            if (!r.pu.empty()) s += "\tPU:" + r.pu;
            if (!r.sm.empty()) s += "\tSM:" + r.sm;
            s += "\n";
        }
        for (const PgRec &r : pg) {
            s += "@PG\tID:" + r.id;
            if (!r.pn.empty()) s += "\tPN:" + r.pn;
            if (!r.cl.empty()) s += "\tCL:" + r.cl;
            if (!r.pp.empty()) s += "\tPP:" + r.pp;
            if (!r.vn.empty()) s += "\tVN:" + r.vn;
            s += "\n";
        }
        for (const std::string &c : co) s += "@CO\t" + c + "\n";
        return s;
    }
};

// CalculateMinimumBin (util/bam_serializer.h:88-98), int arithmetic as there.
inline uint32_t minimum_bin(const int beg, int end) {
    --end;
    if ((beg >> 14) == (end >> 14)) return 4681 + (beg >> 14);
    if ((beg >> 17) == (end >> 17)) return 585 + (beg >> 17);
    if ((beg >> 20) == (end >> 20)) return 73 + (beg >> 20);
    if ((beg >> 23) == (end >> 23)) return 9 + (beg >> 23);
    if ((beg >> 26) == (end >> 26)) return 1 + (beg >> 26);
    return 0;
}

}  // namespace

// ================================================================================================ the file object
struct oge_bam_file {
    oge_bam_alloc_fn alloc_fn = nullptr;
    oge_bam_free_fn free_fn = nullptr;
    uint8_t *stream = nullptr;      // the whole inflated BAM stream (+ slack)
    uint64_t stream_bytes = 0;
    uint64_t first_record = 0;      // stream offset of the first record
    uint64_t rec_bytes = 0;         // bytes of records (shrinks with -r)
    std::vector<uint64_t> offsets;  // n + 1, relative to first_record
    std::string text;
    std::vector<std::pair<std::string, int32_t>> refs;
    HeaderModel header;
    // library table
    std::vector<std::string> rg_ids;
    std::vector<const char *> rg_id_ptrs;
    std::vector<int16_t> rg_libs;
    int16_t unknown_lib = 1;
    int32_t n_libs = 1;
    double t[6] = {0, 0, 0, 0, 0, 0};
    // two-stage open (oge_bam_open_bgzf): the compressed file and its block index, until the records are framed
    uint8_t *comp = nullptr;
    uint64_t comp_bytes = 0;
    BlockIndex ix;
};

namespace {

constexpr int NEED_MORE = 1;      // parse_stream_header: the bytes seen so far end inside the header

int read_whole_file(const char *path, int threads, uint8_t **out, size_t *out_n) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(OGE_BAM_ERR_IO, "cannot open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return fail(OGE_BAM_ERR_IO, "cannot stat %s", path); }
    const size_t fsize = (size_t) st.st_size;
    uint8_t *comp = (uint8_t *) malloc(fsize + 64);
    if (!comp) { close(fd); return fail(OGE_BAM_ERR_NOMEM, "cannot allocate %zu bytes for %s", fsize, path); }
    memset(comp + fsize, 0, 64);
    // parallel pread: one range per worker
    const int rt = (int) std::min<size_t>(threads, std::max<size_t>(1, fsize >> 26));
    const size_t chunk = (fsize + rt - 1) / rt;
    const int rc = parallel_run(rt, [&](int w) -> int {
        size_t at = (size_t) w * chunk;
        const size_t end = std::min(fsize, at + chunk);
        while (at < end) {
            const ssize_t r = pread(fd, comp + at, end - at, (off_t) at);
            if (r <= 0) return fail(OGE_BAM_ERR_IO, "read error on %s", path);
            at += (size_t) r;
        }
        return 0;
    });
    close(fd);
    if (rc) { free(comp); return rc; }
    *out = comp;
    *out_n = fsize;
    return 0;
}

// BamDeserializer::open (util/bam_deserializer.h:40-135) over the first n bytes of the inflated stream.
// complete = these are ALL the bytes of the stream; otherwise running out of bytes returns NEED_MORE.
int parse_stream_header(oge_bam_file *f, const uint8_t *s, uint64_t n, bool complete) {
    auto short_of = [&](const char *msg) { return complete ? fail(OGE_BAM_ERR_FORMAT, "%s", msg) : NEED_MORE; };
    if (n < 12) return short_of("Error reading BAM stream header magic bytes.");
    if (memcmp(s, "BAM\1", 4) != 0) return fail(OGE_BAM_ERR_FORMAT, "Error reading BAM stream header magic bytes.");
    const int32_t l_text = rd_i32(s + 4);
    if (l_text < 0) return fail(OGE_BAM_ERR_FORMAT, "Error reading BAM stream header text.");
    if (8ull + (uint64_t) l_text + 4 > n) return short_of("Error reading BAM stream header text.");
    f->text.assign((const char *) s + 8, (size_t) l_text);
    {   // the reference builds the header from a C string: it ends at the first NUL
        const size_t z = f->text.find('\0');
        if (z != std::string::npos) f->text.resize(z);
    }
    f->header = HeaderModel();
    f->refs.clear();
    int rc = f->header.parse(f->text);
    if (rc) return rc;
    uint64_t pos = 8 + (uint64_t) l_text;
    const int32_t n_ref = rd_i32(s + pos);
    pos += 4;
    if (n_ref < 0) return fail(OGE_BAM_ERR_FORMAT, "Error reading BAM stream reference count.");
    if ((size_t) n_ref != f->header.sq.size())
        return fail(OGE_BAM_ERR_FORMAT, "BAM header text sequence data count doesn't match reference sequence list. Is this file corrupted?");
    for (int32_t i = 0; i < n_ref; i++) {
        if (pos + 4 > n) return short_of("Error reading BAM stream reference sequence name length.");
        const int32_t l_name = rd_i32(s + pos);
        pos += 4;
        if (l_name < 1) return fail(OGE_BAM_ERR_FORMAT, "Error reading BAM stream reference sequence.");
        if (pos + (uint64_t) l_name + 4 > n) return short_of("Error reading BAM stream reference sequence.");
        std::string name((const char *) s + pos, (size_t) l_name - 1);
        pos += (uint64_t) l_name;
        const int32_t len = rd_i32(s + pos);
        pos += 4;
        if (name != f->header.sq[i].name || (long long) len != f->header.sq[i].length)      // :127-131
            return fail(OGE_BAM_ERR_FORMAT, "BAM header text doesn't match sequence information. Is this file corrupted?");
        f->refs.emplace_back(name, len);
    }
    f->first_record = pos;
    return 0;
}

// The record chain: BamDeserializer::read (util/bam_deserializer.h:144-172) over stream[first_record, stream_bytes).
int frame_chain(oge_bam_file *f) {
    const uint8_t *s = f->stream;
    const uint64_t n = f->stream_bytes, base = f->first_record;
    uint64_t pos = base;
    f->offsets.reserve((size_t) ((n - pos) / 160 + 16));
    while (pos < n) {
        if (pos + 4 > n) return fail(OGE_BAM_ERR_FORMAT, "Expected more bytes reading BAM core. Is this file truncated or corrupted?");
        __builtin_prefetch(s + pos + 2048);      // the walk is a dependent chain with a stride the hardware prefetcher cannot guess
        const uint32_t bs = rd_u32(s + pos);
        if (bs < 32 || bs > 10000) return fail(OGE_BAM_ERR_FORMAT, "Invalid BAM block size(%u).", bs);
        if (pos + 4 + bs > n) return fail(OGE_BAM_ERR_FORMAT, "Expected more bytes reading BAM core. Is this file truncated or corrupted?");
        f->offsets.push_back(pos - base);
        pos += 4 + (uint64_t) bs;
    }
    f->offsets.push_back(pos - base);
    f->rec_bytes = pos - base;
    return 0;
}

// library ids (what mark_duplicates.cpp:282-318 resolves per read; only equality of ids matters)
void build_library_table(oge_bam_file *f) {
    std::map<std::string, int16_t> ids;
    int16_t next_id = 1;
    ids["Unknown Library"] = next_id++;
    f->rg_ids.clear();
    f->rg_libs.clear();
    f->rg_id_ptrs.clear();
    for (const RgRec &r : f->header.rg) {
        const std::string lib = r.lb.empty() ? std::string("Unknown Library") : r.lb;
        if (!ids.count(lib)) ids[lib] = next_id++;
        f->rg_ids.push_back(r.id);
        f->rg_libs.push_back(ids[lib]);
    }
    for (const std::string &id : f->rg_ids) f->rg_id_ptrs.push_back(id.c_str());
    f->unknown_lib = 1;
    f->n_libs = (int32_t) ids.size();
}

}  // namespace

extern "C" {

const char *oge_bam_last_error(void) { return g_err; }

void oge_bam_buffer_free(void *p) { free(p); }

void oge_bam_close(oge_bam_file *f) {
    if (!f) return;
    if (f->stream) (f->free_fn ? f->free_fn : free)(f->stream);
    free(f->comp);
    delete f;
}

int oge_bam_load(const char *path, int threads, oge_bam_alloc_fn alloc_fn, oge_bam_free_fn free_fn, oge_bam_file **out) {
    if (!path || !out) return fail(OGE_BAM_ERR_ARG, "load: null argument");
    if ((alloc_fn == nullptr) != (free_fn == nullptr)) return fail(OGE_BAM_ERR_ARG, "load: alloc_fn and free_fn go together");
    threads = clamp_threads(threads);
    *out = nullptr;
    double t0 = now_s();
    uint8_t *comp = nullptr;
    size_t fsize = 0;
    {
        const int rc = read_whole_file(path, threads, &comp, &fsize);
        if (rc) return rc;
    }
    oge_bam_file *f = new oge_bam_file();
    f->alloc_fn = alloc_fn;
    f->free_fn = free_fn;
    f->t[0] = now_s() - t0;
    auto bail = [&](int rc) {
        free(comp);
        oge_bam_close(f);
        return rc;
    };
    auto alloc = [&](size_t n) { return (uint8_t *) (alloc_fn ? alloc_fn(n) : malloc(n)); };

    if (fsize >= 4 && memcmp(comp, "BAM\1", 4) == 0) {      // an uncompressed stream ("rawbam")
        f->stream = alloc(fsize + 256);
        if (!f->stream) return bail(fail(OGE_BAM_ERR_NOMEM, "cannot allocate %zu bytes", fsize + 256));
        memcpy(f->stream, comp, fsize);
        f->stream_bytes = fsize;
    } else {
        t0 = now_s();
        BlockIndex ix;
        int rc = bgzf_scan(comp, fsize, ix);
        if (rc) return bail(rc);
        f->t[1] = now_s() - t0;
        t0 = now_s();
        f->stream_bytes = ix.out_off.back();
        f->stream = alloc(f->stream_bytes + 256);
        if (!f->stream) return bail(fail(OGE_BAM_ERR_NOMEM, "cannot allocate %llu bytes", (unsigned long long) f->stream_bytes + 256));
        rc = bgzf_inflate(comp, ix, f->stream, threads);
        if (rc) return bail(rc);
        f->t[2] = now_s() - t0;
    }
    free(comp);
    comp = nullptr;
    memset(f->stream + f->stream_bytes, 0, 256);

    t0 = now_s();
    int rc = parse_stream_header(f, f->stream, f->stream_bytes, true);
    if (rc) return bail(rc);
    rc = frame_chain(f);
    if (rc) return bail(rc);
    build_library_table(f);
    f->t[3] = now_s() - t0;
    *out = f;
    return 0;
}

// ---- two-stage open for callers that inflate elsewhere (the GPU: oge_gpu_dedup_push_bgzf) -------------------------
int oge_bam_open_bgzf(const char *path, int threads, oge_bam_alloc_fn alloc_fn, oge_bam_free_fn free_fn, oge_bam_file **out) {
    if (!path || !out) return fail(OGE_BAM_ERR_ARG, "open_bgzf: null argument");
    if ((alloc_fn == nullptr) != (free_fn == nullptr)) return fail(OGE_BAM_ERR_ARG, "open_bgzf: alloc_fn and free_fn go together");
    threads = clamp_threads(threads);
    *out = nullptr;
    double t0 = now_s();
    uint8_t *comp = nullptr;
    size_t fsize = 0;
    int rc = read_whole_file(path, threads, &comp, &fsize);
    if (rc) return rc;
    oge_bam_file *f = new oge_bam_file();
    f->alloc_fn = alloc_fn;
    f->free_fn = free_fn;
    f->comp = comp;
    f->comp_bytes = fsize;
    f->t[0] = now_s() - t0;
    auto bail = [&](int code) {
        oge_bam_close(f);
        return code;
    };
    if (fsize >= 4 && memcmp(comp, "BAM\1", 4) == 0) return bail(fail(OGE_BAM_ERR_FORMAT, "open_bgzf: %s is an uncompressed BAM stream (use oge_bam_load)", path));
    t0 = now_s();
    if ((rc = bgzf_scan(comp, fsize, f->ix))) return bail(rc);
    f->t[1] = now_s() - t0;
    f->stream_bytes = f->ix.out_off.back();
    // the header decides where the records start: inflate leading blocks on the host until it is complete
    t0 = now_s();
    std::vector<uint8_t> head;
    size_t nb = 0;
    while (true) {
        rc = parse_stream_header(f, head.data(), head.size(), nb == f->ix.in_off.size());
        if (rc == 0) break;
        if (rc != NEED_MORE) return bail(rc);
        // next block
        const size_t b = nb++;
        const size_t at = head.size();
        head.resize(at + f->ix.isize[b] + 8);
        BlockIndex one;
        one.in_off.push_back(f->ix.in_off[b]);
        one.csize.push_back(f->ix.csize[b]);
        one.isize.push_back(f->ix.isize[b]);
        one.out_off.push_back(0);
        one.out_off.push_back(f->ix.isize[b]);
        if ((rc = bgzf_inflate(comp, one, head.data() + at, 1))) return bail(rc);
        head.resize(at + f->ix.isize[b]);
    }
    f->rec_bytes = f->stream_bytes - f->first_record;
    build_library_table(f);
    f->t[3] = now_s() - t0;
    *out = f;
    return 0;
}

int oge_bam_bgzf_index(const oge_bam_file *f, const uint8_t **comp, uint64_t *comp_bytes, const uint64_t **block_in_off,
                       const uint32_t **block_csize, const uint32_t **block_isize, uint64_t *n_blocks, uint64_t *header_bytes) {
    if (!f || !f->comp) return fail(OGE_BAM_ERR_ARG, "bgzf_index: the file was not opened with oge_bam_open_bgzf");
    if (comp) *comp = f->comp;
    if (comp_bytes) *comp_bytes = f->comp_bytes;
    if (block_in_off) *block_in_off = f->ix.in_off.data();
    if (block_csize) *block_csize = f->ix.csize.data();
    if (block_isize) *block_isize = f->ix.isize.data();
    if (n_blocks) *n_blocks = f->ix.in_off.size();
    if (header_bytes) *header_bytes = f->first_record;
    return 0;
}

uint8_t *oge_bam_records_buffer(oge_bam_file *f) {
    if (!f || (!f->comp && !f->stream)) { fail(OGE_BAM_ERR_ARG, "records_buffer: the file was not opened with oge_bam_open_bgzf"); return nullptr; }
    if (!f->stream) {
        // laid out like a loaded file (header region left unused) so that every other entry point works unchanged
        f->stream = (uint8_t *) (f->alloc_fn ? f->alloc_fn(f->stream_bytes + 256) : malloc(f->stream_bytes + 256));
        if (!f->stream) { fail(OGE_BAM_ERR_NOMEM, "cannot allocate %llu bytes", (unsigned long long) f->stream_bytes + 256); return nullptr; }
        if (!f->alloc_fn) {
            // fresh pageable memory: fault the pages in with all threads now (callers overlap this with the GPU's work),
            // not one by one under the device-to-host copy
            const uint64_t total = f->stream_bytes + 256;
            const int th = clamp_threads(0);
            const uint64_t per = (total / th + 4096) & ~4095ull;
            uint8_t *base = f->stream;
            parallel_run(th, [&](int w) -> int {
                for (uint64_t at = (uint64_t) w * per; at < std::min(total, (uint64_t) (w + 1) * per); at += 4096) base[at] = 0;
                return 0;
            });
        }
    }
    return f->stream + f->first_record;
}

int oge_bam_adopt_offsets(oge_bam_file *f, const uint64_t *offsets, uint64_t nrec) {
    if (!f || !offsets || !f->stream) return fail(OGE_BAM_ERR_ARG, "adopt_offsets: fill oge_bam_records_buffer first");
    if (offsets[0] != 0 || offsets[nrec] != f->stream_bytes - f->first_record) return fail(OGE_BAM_ERR_ARG, "adopt_offsets: offsets do not span the record bytes");
    const double t0 = now_s();
    memset(f->stream + f->stream_bytes, 0, 256);
    free(f->comp);
    f->comp = nullptr;
    f->offsets.assign(offsets, offsets + nrec + 1);
    f->rec_bytes = offsets[nrec];
    f->t[3] += now_s() - t0;
    return 0;
}

int oge_bam_set_sort_order(oge_bam_file *f, const char *so) {
    if (!f || !so) return fail(OGE_BAM_ERR_ARG, "set_sort_order: null argument");
    static const char *names[] = {"unknown", "unsorted", "queryname", "coordinate"};
    for (int i = 0; i < 4; i++)
        if (!strcmp(so, names[i])) {
            f->header.sort = i;
            if (f->header.version.empty()) f->header.version = "1.4";
            return 0;
        }
    return fail(OGE_BAM_ERR_ARG, "Unknown sort order '%s'.", so);
}

int oge_bam_frame_records(oge_bam_file *f) {
    if (!f || !f->comp || !f->stream) return fail(OGE_BAM_ERR_ARG, "frame_records: call oge_bam_open_bgzf and fill oge_bam_records_buffer first");
    const double t0 = now_s();
    memset(f->stream + f->stream_bytes, 0, 256);
    free(f->comp);      // the compressed bytes are not needed any more
    f->comp = nullptr;
    f->offsets.clear();
    const int rc = frame_chain(f);
    f->t[3] += now_s() - t0;
    return rc;
}

const char *oge_bam_header_text(const oge_bam_file *f) { return f ? f->text.c_str() : ""; }
int32_t oge_bam_n_ref(const oge_bam_file *f) { return f ? (int32_t) f->refs.size() : 0; }
const char *oge_bam_ref_name(const oge_bam_file *f, int32_t i) { return f && i >= 0 && (size_t) i < f->refs.size() ? f->refs[i].first.c_str() : ""; }
int32_t oge_bam_ref_len(const oge_bam_file *f, int32_t i) { return f && i >= 0 && (size_t) i < f->refs.size() ? f->refs[i].second : 0; }
uint8_t *oge_bam_records(oge_bam_file *f) { return f ? f->stream + f->first_record : nullptr; }
uint64_t oge_bam_records_bytes(const oge_bam_file *f) { return f ? f->rec_bytes : 0; }
const uint64_t *oge_bam_offsets(const oge_bam_file *f) { return f ? f->offsets.data() : nullptr; }
uint64_t oge_bam_n_records(const oge_bam_file *f) { return f && !f->offsets.empty() ? f->offsets.size() - 1 : 0; }

int oge_bam_library_table(oge_bam_file *f, const char *const **ids, const int16_t **lib_ids, int32_t *n, int16_t *unknown_lib_id,
                          int32_t *n_libs) {
    if (!f || !ids || !lib_ids || !n || !unknown_lib_id || !n_libs) return fail(OGE_BAM_ERR_ARG, "library_table: null argument");
    *ids = f->rg_id_ptrs.empty() ? nullptr : f->rg_id_ptrs.data();
    *lib_ids = f->rg_libs.empty() ? nullptr : f->rg_libs.data();
    *n = (int32_t) f->rg_ids.size();
    *unknown_lib_id = f->unknown_lib;
    *n_libs = f->n_libs;
    return 0;
}

int oge_bam_apply_flags(oge_bam_file *f, const uint16_t *flags, int remove_duplicates, int threads) {
    if (!f || (!flags && oge_bam_n_records(f))) return fail(OGE_BAM_ERR_ARG, "apply_flags: null argument");
    threads = clamp_threads(threads);
    const double t0 = now_s();
    if (f->offsets.empty()) return fail(OGE_BAM_ERR_ARG, "apply_flags: the records have not been framed yet");
    const uint64_t n = f->offsets.size() - 1;
    uint8_t *rec = f->stream + f->first_record;
    const uint64_t *off = f->offsets.data();
    std::atomic<uint64_t> next(0);
    const uint64_t CH = 1 << 16;
    int rc = parallel_run((int) std::min<uint64_t>(threads, n / CH + 1), [&](int) -> int {
        while (true) {
            const uint64_t i0 = next.fetch_add(CH);
            if (i0 >= n) return 0;
            for (uint64_t i = i0; i < std::min(n, i0 + CH); i++) {
                uint8_t *p = rec + off[i];
                const uint32_t bs = rd_u32(p), l_name = p[12], n_cig = rd_u16(p + 16);
                if (36ull + l_name + 4ull * n_cig > 4ull + bs) return fail(OGE_BAM_ERR_FORMAT, "record %llu: name and CIGAR overrun the record", (unsigned long long) i);
                // the writer's bin: CalculateMinimumBin(pos, GetEndPosition()) (bam_serializer.h:112-116,
                // bamtools/BamAlignment.cpp:311-350: pos + lengths of M D N = X)
                const int32_t pos = rd_i32(p + 8);
                int32_t end = pos;
                const uint8_t *cg = p + 36 + l_name;
                for (uint32_t k = 0; k < n_cig; k++) {
                    const uint32_t c = rd_u32(cg + 4 * k), op = c & 0xF;
                    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += (int32_t) (c >> 4);
                }
                wr_u16(p + 14, (uint16_t) minimum_bin(pos, end));
                wr_u16(p + 18, flags[i]);
            }
        }
    });
    if (rc) return rc;
    if (remove_duplicates) {      // mark_duplicates.cpp:456-458: anything flagged after the rewrite is dropped
        uint64_t w = 0, kept = 0;
        for (uint64_t i = 0; i < n; i++) {
            const uint64_t a = off[i], len = off[i + 1] - a;
            if (flags[i] & 0x400) continue;
            if (w != a) memmove(rec + w, rec + a, len);
            f->offsets[kept++] = w;
            w += len;
        }
        f->offsets[kept] = w;
        f->offsets.resize(kept + 1);
        f->rec_bytes = w;
    }
    f->t[4] = now_s() - t0;
    return 0;
}

// header bytes: BamSerializer::open (util/bam_serializer.h:46-79) over the re-rendered header
static std::vector<uint8_t> render_head(const oge_bam_file *f, const char *pg_command_line, const char *pg_version) {
    HeaderModel h = f->header;
    if (pg_command_line && *pg_command_line) {      // file_writer.cpp:76-89
        PgRec pg;
        pg.id = "openge";
        pg.vn = pg_version ? pg_version : "";
        for (int i = 2;; i++) {
            bool taken = false;
            for (const PgRec &r : h.pg) taken = taken || r.id == pg.id;
            if (!taken) break;
            pg.id = "openge-" + std::to_string(i);
        }
        pg.cl = pg_command_line;
        h.pg.push_back(pg);
    }
    const std::string text = h.render();
    std::vector<uint8_t> head;
    auto put = [&](const void *p, size_t n) { head.insert(head.end(), (const uint8_t *) p, (const uint8_t *) p + n); };
    put("BAM\1", 4);
    const int32_t l_text = (int32_t) text.size();
    put(&l_text, 4);
    put(text.data(), text.size());
    const int32_t n_ref = (int32_t) h.sq.size();
    put(&n_ref, 4);
    for (const SqRec &r : h.sq) {
        const int32_t l_name = (int32_t) r.name.size() + 1, len = (int32_t) r.length;
        put(&l_name, 4);
        put(r.name.c_str(), r.name.size() + 1);
        put(&len, 4);
    }

    return head;
}

int oge_bam_store(oge_bam_file *f, const char *path, const char *format, int level, const char *pg_command_line,
                  const char *pg_version, int threads) {
    if (!f || !path) return fail(OGE_BAM_ERR_ARG, "store: null argument");
    threads = clamp_threads(threads);
    const double t0 = now_s();
    bool raw = false;
    if (format && *format) {
        if (!strcmp(format, "rawbam")) raw = true;
        else if (strcmp(format, "bam")) return fail(OGE_BAM_ERR_ARG, "Unknown file format specified: %s.", format);
    }
    if (level < 0 || level > 9) return fail(OGE_BAM_ERR_ARG, "store: compression level %d", level);

    const std::vector<uint8_t> head = render_head(f, pg_command_line, pg_version);

    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return fail(OGE_BAM_ERR_IO, "cannot open %s for writing: %s", path, strerror(errno));
    Sink sink;
    sink.fd = fd;
    int rc = 0;
    if (raw) {
        rc = sink.write(head.data(), head.size());
        if (!rc) rc = sink.write(f->stream + f->first_record, f->rec_bytes);
    } else {
        Segments seg;
        seg.add(head.data(), head.size());
        seg.add(f->stream + f->first_record, f->rec_bytes);
        rc = bgzf_deflate_stream(seg, level, threads, sink);
    }
    if (close(fd) != 0 && !rc) rc = fail(OGE_BAM_ERR_IO, "close failed on %s", path);
    f->t[5] = now_s() - t0;
    return rc;
}

int oge_bam_store_members_stream(oge_bam_file *f, const char *path, int level, const char *pg_command_line, const char *pg_version,
                                 oge_bam_fill_fn fill, void *user, uint64_t members_bytes_hint, int threads) {
    if (!f || !path || !fill) return fail(OGE_BAM_ERR_ARG, "store_members: null argument");
    if (level < 0 || level > 9) return fail(OGE_BAM_ERR_ARG, "store_members: compression level %d", level);
    threads = clamp_threads(threads);
    const double t0 = now_s();
    const std::vector<uint8_t> head = render_head(f, pg_command_line, pg_version);
    // the header in members of its own (host zlib, `level`), then the record members as they come, then the empty member that
    // ends a BGZF file (BgzfOutputStream::close, util/bgzf_output_stream.cpp:225-250)
    std::vector<uint8_t> front, hz(BGZF_BLOCK);
    const uint32_t full = level == 0 ? BGZF_BLOCK - 64 : BGZF_BLOCK;
    int rc = 0;
    for (size_t at = 0; at < head.size() && !rc; at += full) {
        uint32_t n = 0;
        rc = bgzf_compress_block(head.data() + at, (uint32_t) std::min<size_t>(full, head.size() - at), level, hz.data(), &n);
        if (!rc) front.insert(front.end(), hz.data(), hz.data() + n);
    }
    uint32_t eof_n = 0;
    if (!rc) rc = bgzf_compress_block(head.data(), 0, level, hz.data(), &eof_n);
    if (rc) return rc;
    const int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return fail(OGE_BAM_ERR_IO, "cannot open %s for writing: %s", path, strerror(errno));
    // With the size of the members known the file is sized first and filled through a shared mapping by several threads: one
    // write() stream into the page cache runs at 3-4 GB/s, and writers at disjoint offsets of one file queue up behind its lock.
    const uint64_t total = front.size() + members_bytes_hint + eof_n;
    uint8_t *map = nullptr;
    if (members_bytes_hint >= ((uint64_t) 64 << 20) && threads > 1 && ftruncate(fd, (off_t) total) == 0) {
        void *m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (m != MAP_FAILED) map = (uint8_t *) m;
        else if (ftruncate(fd, 0) != 0) rc = fail(OGE_BAM_ERR_IO, "truncate failed on %s: %s", path, strerror(errno));
    }
    Sink sink;
    sink.fd = fd;
    uint64_t pos = 0;
    auto put = [&](const uint8_t *p, uint64_t n) -> int {
        if (!map) return sink.write(p, n);
        if (pos + n > total) return fail(OGE_BAM_ERR_ARG, "store_members: more member bytes than announced");
        const int parts = (int) std::max<uint64_t>(1, std::min<uint64_t>((uint64_t) threads, n / ((uint64_t) 4 << 20)));
        const uint64_t per = (n + parts - 1) / parts;
        uint8_t *dst = map + pos;
        if (parts == 1) memcpy(dst, p, n);
        else parallel_run(parts, [&](int w) -> int {
            const uint64_t a = per * w, b = std::min(n, a + per);
            if (b > a) memcpy(dst + a, p + a, b - a);
            return 0;
        });
        pos += n;
        return 0;
    };
    if (!rc) rc = put(front.data(), front.size());
    uint64_t got = 0;
    while (!rc) {
        const uint8_t *data = nullptr;
        uint64_t n = 0;
        if (fill(user, &data, &n)) {
            rc = fail(OGE_BAM_ERR_IO, "store_members: the source of the members failed");
            break;
        }
        if (!n) break;
        got += n;
        rc = put(data, n);
    }
    if (!rc) rc = put(hz.data(), eof_n);
    if (map) {
        if (munmap(map, total) != 0 && !rc) rc = fail(OGE_BAM_ERR_IO, "munmap failed on %s: %s", path, strerror(errno));
        if (!rc && got != members_bytes_hint) rc = fail(OGE_BAM_ERR_ARG, "store_members: %llu member bytes announced, %llu delivered",
                                                        (unsigned long long) members_bytes_hint, (unsigned long long) got);
    }
    if (close(fd) != 0 && !rc) rc = fail(OGE_BAM_ERR_IO, "close failed on %s", path);
    f->t[5] = now_s() - t0;
    return rc;
}

namespace {
struct OneChunk {
    const uint8_t *p;
    uint64_t n;
};
int one_chunk_fill(void *user, const uint8_t **data, uint64_t *nbytes) {
    OneChunk *c = (OneChunk *) user;
    *data = c->p;
    *nbytes = c->n;
    c->n = 0;
    return 0;
}
}  // namespace

int oge_bam_store_members(oge_bam_file *f, const char *path, int level, const char *pg_command_line, const char *pg_version,
                          const uint8_t *members, uint64_t members_bytes) {
    if (!members && members_bytes) return fail(OGE_BAM_ERR_ARG, "store_members: null argument");
    OneChunk c = {members, members_bytes};
    return oge_bam_store_members_stream(f, path, level, pg_command_line, pg_version, one_chunk_fill, &c, members_bytes, 0);
}

int oge_bam_timings(const oge_bam_file *f, double *out, int n) {
    if (!f || !out) return fail(OGE_BAM_ERR_ARG, "timings: null argument");
    for (int i = 0; i < n && i < 6; i++) out[i] = f->t[i];
    return 0;
}

int oge_bgzf_decompress(const uint8_t *in, size_t n, int threads, uint8_t **out, size_t *out_n) {
    if ((!in && n) || !out || !out_n) return fail(OGE_BAM_ERR_ARG, "decompress: null argument");
    BlockIndex ix;
    int rc = bgzf_scan(in, n, ix);
    if (rc) return rc;
    const size_t total = (size_t) ix.out_off.back();
    uint8_t *buf = (uint8_t *) malloc(total + 1);
    if (!buf) return fail(OGE_BAM_ERR_NOMEM, "cannot allocate %zu bytes", total);
    rc = bgzf_inflate(in, ix, buf, clamp_threads(threads));
    if (rc) { free(buf); return rc; }
    *out = buf;
    *out_n = total;
    return 0;
}

int oge_bgzf_compress(const uint8_t *in, size_t n, int level, int threads, uint8_t **out, size_t *out_n) {
    if ((!in && n) || !out || !out_n) return fail(OGE_BAM_ERR_ARG, "compress: null argument");
    if (level < 0 || level > 9) return fail(OGE_BAM_ERR_ARG, "compress: level %d", level);
    std::vector<uint8_t> buf;
    Sink sink;
    sink.buf = &buf;
    Segments seg;
    seg.add(in, n);
    const int rc = bgzf_deflate_stream(seg, level, clamp_threads(threads), sink);
    if (rc) return rc;
    uint8_t *p = (uint8_t *) malloc(buf.size() + 1);
    if (!p) return fail(OGE_BAM_ERR_NOMEM, "cannot allocate %zu bytes", buf.size());
    memcpy(p, buf.data(), buf.size());
    *out = p;
    *out_n = buf.size();
    return 0;
}

int oge_bam_header_render(const char *text, char **out) {
    if (!text || !out) return fail(OGE_BAM_ERR_ARG, "header_render: null argument");
    HeaderModel h;
    const int rc = h.parse(text);
    if (rc) return rc;
    const std::string s = h.render();
    char *p = (char *) malloc(s.size() + 1);
    if (!p) return fail(OGE_BAM_ERR_NOMEM, "cannot allocate %zu bytes", s.size());
    memcpy(p, s.c_str(), s.size() + 1);
    *out = p;
    return 0;
}

}  // extern "C"
