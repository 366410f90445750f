// Drop-in replacement for the reference's src/algorithms/read_sorter.cpp: the SAME class against the reference's own
// unmodified header (src/algorithms/read_sorter.h), so that `openge mergesort` (commands/command_mergesort.cpp:68-100)
// and every other chain with a ReadSorter in it compile and run unchanged, with the coordinate sort on the GPU.
//
//   reference (read_sorter.cpp)                               here
//   GenerateSortedRuns (:119-192): runs of 200 000 reads       the reads are framed into pinned batches and pushed to the
//     sorted with ogeSortMt(..., Sort::ByPosition()) (:203-      device as they fill
//     205) and spilled as temp BAM files
//   MergeSortedRuns (:66-117): a std::multiset of the same     oge_gpu_dedup_sort: one device sort of everything (csrc/
//     comparator over the temp files' heads                      coordsort.cu; the order of Sort::ByPosition, util/bamtools/
//                                                                Sort.h:108-133 -- where the reference falls through to the
//                                                                ADDRESSES of its heap objects, the input order is kept)
//   putOutputAlignment in merged order                         oge_gpu_dedup_pull: the records in sorted order; reads
//                                                                rebuilt from them and passed on
//
// Only coordinate order is built (the default of `openge mergesort`; --byname needs Sort::ByName).  Nothing is written to
// the temp directory.  Compiled as gnu++98 like the rest of the reference (openge_b200/host/Makefile).
#include "algorithms/read_sorter.h"

#include <unistd.h>

#include "oge_gpu_dedup.h"
#include "record_batch.h"

using namespace std;
using namespace oge_host;

namespace {

void sorter_fail(const char * what, int rc) {
    cerr << "ReadSorter (GPU): " << what << " failed (" << rc << "): " << oge_gpu_last_error() << endl;
    exit(-1);
}

int sorter_env_int(const char * name, int dflt) {
    const char * v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

}  // namespace

const BamHeader & ReadSorter::getHeader()      // read_sorter.cpp:233-246: downstream modules wait for the sorter's header
{
    while (true) {
        m_header_access.lock();
        bool ret = header_loaded;
        m_header_access.unlock();
        if (ret) break;
        usleep(10000);
    }
    return m_header;
}

int ReadSorter::runInternal()
{
    ogeNameThread("am_ReadSorter");
    m_header_access.lock();
    m_header = AlgorithmModule::getHeader();
    m_header.setSortOrder(sort_order);      // :256-258
    header_loaded = true;
    m_header_access.unlock();
    if (sort_order != BamHeader::SORT_COORDINATE) {
        cerr << "ReadSorter (GPU): only coordinate order is built on the device. Aborting." << endl;
        exit(-1);
    }

    oge_gpu_dedup_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = OGE_GPU_DEDUP_ABI_VERSION;
    cfg.device = sorter_env_int("OGE_GPU_DEVICE", 0);
    cfg.n_ref = (int32_t) m_header.getSequences().size();
    for (size_t i = 0; i < m_header.getSequences().size(); i++)
        if (m_header.getSequences()[(int) i].getLength() > cfg.max_ref_len) cfg.max_ref_len = m_header.getSequences()[(int) i].getLength();
    cfg.verify_names = -1;
    oge_gpu_dedup_ctx * ctx = NULL;
    int rc = oge_gpu_dedup_create(&cfg, &ctx);
    if (rc) sorter_fail("oge_gpu_dedup_create", rc);

    // ---- the reads, framed into pinned batches and pushed as they fill (GenerateSortedRuns' read loop, :131-160)
    vector<Batch *> batches;
    Batch * cur = NULL;
    uint64_t total_bytes = 0;
    m_numberOfAlignments = 0;
    while (true) {
        OGERead * al = getInputAlignment();
        if (!al) break;
        const size_t rec_len = record_bytes(*al);
        if (rec_len > BATCH_BYTES) { cerr << "ReadSorter (GPU): record of " << rec_len << " bytes. Aborting." << endl; exit(-1); }
        if (!cur || cur->used + rec_len > BATCH_BYTES) {
            if (cur) {
                rc = oge_gpu_dedup_push(ctx, cur->data, cur->used, &cur->offsets[0], cur->offsets.size() - 1);
                if (rc) sorter_fail("oge_gpu_dedup_push", rc);
            }
            cur = new Batch();
            cur->data = (uint8_t *) oge_gpu_host_alloc(BATCH_BYTES);
            if (!cur->data) { cerr << "ReadSorter (GPU): cannot allocate a pinned staging buffer. Aborting." << endl; exit(-1); }
            batches.push_back(cur);
        }
        append_read(*cur, *al);
        total_bytes += rec_len;
        OGERead::deallocate(al);
        m_numberOfAlignments++;
    }
    if (cur && cur->offsets.size() > 1) {
        rc = oge_gpu_dedup_push(ctx, cur->data, cur->used, &cur->offsets[0], cur->offsets.size() - 1);
        if (rc) sorter_fail("oge_gpu_dedup_push", rc);
    }
    if (isVerbose()) cerr << "Sorting " << m_numberOfAlignments << " reads on the GPU." << endl;

    // ---- ReadSorter's whole run / merge machinery: one device sort
    rc = oge_gpu_dedup_sort(ctx);
    if (rc) sorter_fail("oge_gpu_dedup_sort", rc);
    if (isVerbose()) {
        uint64_t tied = 0, rounds = 0, launches = 0;
        float ms = 0;
        oge_gpu_dedup_sort_stats(ctx, &tied, &rounds, &launches, &ms);
        cerr << "Sorted " << m_numberOfAlignments << " records by coordinate on the GPU in " << ms << " ms (" << tied << " tied on position, "
             << rounds << " name rounds, " << launches << " kernel launches)." << endl;
    }
    // the staging batches are free again: the sorted records come back into them, batch by batch would need the offsets
    // first, so one buffer takes them all
    for (size_t b = 0; b < batches.size(); b++) {
        oge_gpu_host_free(batches[b]->data);
        delete batches[b];
    }
    const uint64_t n = (uint64_t) m_numberOfAlignments;
    // one copy back: page-locking the whole output (0.4 s per GB) would cost more than the pageable copy loses
    uint8_t * sorted = (uint8_t *) malloc(total_bytes ? total_bytes : 1);
    vector<uint64_t> offs(n + 1);
    if (!sorted) { cerr << "ReadSorter (GPU): cannot allocate the output buffer. Aborting." << endl; exit(-1); }
    uint64_t got_bytes = 0, got_n = 0;
    rc = oge_gpu_dedup_pull(ctx, sorted, total_bytes, &offs[0], n + 1, &got_bytes, &got_n);
    if (rc) sorter_fail("oge_gpu_dedup_pull", rc);
    oge_gpu_dedup_destroy(ctx);

    // ---- MergeSortedRuns' output loop (:100-112)
    for (uint64_t i = 0; i < got_n; i++) {
        const uint8_t * p = sorted + offs[i];
        putOutputAlignment(rebuild_read(p, (uint16_t) (get_u32(p + 16) >> 16)));
    }
    free(sorted);
    return true;      // the reference returns its bool (read_sorter.cpp:272)
}

// The reference's class declares a nested job type with a virtual method; nothing here spills temp files, but the
// vtable wants its definitions.
ReadSorter::TempFileWriteJob::TempFileWriteJob(ReadSorter * tool, vector<OGERead *> * buffer, string filename)
: filename(filename), buffer(buffer), tool(tool) {}
void ReadSorter::TempFileWriteJob::runJob() {}
