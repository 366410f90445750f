// Drop-in replacement for the reference's src/algorithms/mark_duplicates.cpp.
//
// It implements the SAME class, against the reference's own unmodified header
// (src/algorithms/mark_duplicates.h:27-68): constructor MarkDuplicates(std::string temp_directory),
// public field removeDuplicates, and the AlgorithmModule entry point runInternal().  Link this file
// instead of mark_duplicates.cpp and `openge dedup` (commands/command_dedup.cpp:37-114) and
// `openge mergesort -M/-R` (commands/command_mergesort.cpp:80-100) compile and run unchanged; the
// work happens on the GPU through the C ABI of include/oge_gpu_dedup.h.
//
// What runInternal() does, next to the reference (mark_duplicates.cpp:422-475):
//   reference                                         here
//   getInputAlignment() loop, temp-file spill         records are framed (bam_serializer.h:106-141 layout) into
//     (:192-256)                                        pinned host batches and pushed to the device as they fill
//   buildSortedReadEndLists + generateDuplicate-      oge_gpu_dedup_run()
//     Indexes (:185-279, :326-400)
//   re-read the spill, SetIsDuplicate, -r filter,     flags come back (u16 per record); records are rebuilt from
//     putOutputAlignment (:435-465)                      the host batches with the new flag word and passed on
//
// Built only where the reference sources are available (openge_b200/host/Makefile); compiled as
// gnu++98 like the rest of the reference.
#include "algorithms/mark_duplicates.h"

#include "oge_gpu_dedup.h"
#include "record_batch.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <iostream>
#include <map>
#include <string>
#include <vector>

using namespace std;
using namespace oge_host;

namespace {

void gpu_fail(const char * what, int rc) {
    // the reference's convention for fatal errors: message on cerr, exit(-1) (e.g. util/bam_deserializer.h:155-163)
    cerr << "MarkDuplicates (GPU): " << what << " failed (" << rc << "): " << oge_gpu_last_error() << endl;
    exit(-1);
}

int env_int(const char * name, int dflt) {
    const char * v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

}  // namespace

MarkDuplicates::MarkDuplicates(string temp_directory)
: numDuplicateIndices(0)
, nextLibraryId(1)
, bufferFilename(temp_directory)    // kept for interface parity: nothing is spilled to disk
, removeDuplicates(false)
{
}

int MarkDuplicates::runInternal() {
    ogeNameThread("am_MarkDuplicates");

    BamHeader header = getHeader();

    // ---- context sized from the reference dictionary
    oge_gpu_dedup_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = OGE_GPU_DEDUP_ABI_VERSION;
    cfg.device = env_int("OGE_GPU_DEVICE", 0);
    cfg.n_ref = (int32_t) header.getSequences().size();
    for (size_t i = 0; i < header.getSequences().size(); i++)
        if (header.getSequences()[(int) i].getLength() > cfg.max_ref_len) cfg.max_ref_len = header.getSequences()[(int) i].getLength();
    cfg.remove_duplicates = removeDuplicates ? 1 : 0;
    cfg.verify_names = -1;
    // without -v the reference never advances its record index.  That is a bug, not a
    // feature, so it is reproduced only on request.
    cfg.compat_quiet_index_bug = env_int("OGE_COMPAT_QUIET_INDEX_BUG", 0) && !verbose;
    oge_gpu_dedup_ctx * ctx = NULL;
    int rc = oge_gpu_dedup_create(&cfg, &ctx);
    if (rc) gpu_fail("oge_gpu_dedup_create", rc);

    // ---- @RG ID -> LB -> library id, resolved once (getLibraryName/getLibraryId, mark_duplicates.cpp:282-318)
    {
        static const string unknown_library("Unknown Library");
        const BamReadGroupRecords & rgs = header.getReadGroups();
        vector<const char *> ids;
        vector<int16_t> libs;
        libraryIds[unknown_library] = nextLibraryId++;
        for (BamReadGroupRecords::const_iterator g = rgs.begin(); g != rgs.end(); g++) {
            const string & name = g->getLibrary().empty() ? unknown_library : g->getLibrary();
            if (!libraryIds.count(name)) libraryIds[name] = nextLibraryId++;
            ids.push_back(g->getId().c_str());      // a repeated ID never matches: the first one wins, as in BamReadGroupRecords::operator[]
            libs.push_back(libraryIds[name]);
        }
        rc = oge_gpu_dedup_set_readgroups(ctx, ids.empty() ? NULL : &ids[0], libs.empty() ? NULL : &libs[0], (int32_t) ids.size(),
                                          libraryIds[unknown_library], (int32_t) libraryIds.size());
        if (rc) gpu_fail("oge_gpu_dedup_set_readgroups", rc);
    }

    if (verbose) cerr << "Reading input file and constructing read end information." << endl;

    // ---- pass 1: frame the records into pinned batches, push each batch as it fills
    vector<Batch *> batches;
    Batch * cur = NULL;
    uint64_t n_records = 0;
    while (true) {
        OGERead * al = getInputAlignment();
        if (!al) break;
        const size_t rec_len = record_bytes(*al);
        if (rec_len > BATCH_BYTES) { cerr << "MarkDuplicates (GPU): record of " << rec_len << " bytes. Aborting." << endl; exit(-1); }
        if (!cur || cur->used + rec_len > BATCH_BYTES) {
            if (cur) {
                rc = oge_gpu_dedup_push(ctx, cur->data, cur->used, &cur->offsets[0], cur->offsets.size() - 1);
                if (rc) gpu_fail("oge_gpu_dedup_push", rc);
            }
            cur = new Batch();
            cur->data = (uint8_t *) oge_gpu_host_alloc(BATCH_BYTES);
            if (!cur->data) { cerr << "MarkDuplicates (GPU): cannot allocate a pinned staging buffer. Aborting." << endl; exit(-1); }
            batches.push_back(cur);
        }
        append_read(*cur, *al);
        OGERead::deallocate(al);
        n_records++;
        if (verbose && n_records % 100000 == 0) cerr << "\rRead " << n_records << " records." << std::flush;
    }
    if (cur && cur->offsets.size() > 1) {
        rc = oge_gpu_dedup_push(ctx, cur->data, cur->used, &cur->offsets[0], cur->offsets.size() - 1);
        if (rc) gpu_fail("oge_gpu_dedup_push", rc);
    }
    if (verbose) cerr << "\rRead " << n_records << " records." << endl;

    // ---- the whole of buildSortedReadEndLists + generateDuplicateIndexes + the flag rewrite
    rc = oge_gpu_dedup_run(ctx);
    if (rc) gpu_fail("oge_gpu_dedup_run", rc);
    oge_gpu_dedup_stats st;
    oge_gpu_dedup_get_stats(ctx, &st);
    numDuplicateIndices = (int) st.n_duplicates;
    if (verbose) {
        cerr << "Sorted " << st.n_pair_entries << " pair ends and " << st.n_frag_entries << " fragment ends on the GPU in "
             << st.ms_total << " ms (" << st.launches << " kernel launches)." << endl;
        cerr << "Marking " << numDuplicateIndices << " records as duplicates." << endl;
    }
    vector<uint16_t> flags(n_records ? n_records : 1);
    rc = oge_gpu_dedup_flags(ctx, &flags[0], n_records);
    if (rc) gpu_fail("oge_gpu_dedup_flags", rc);
    oge_gpu_dedup_destroy(ctx);

    // ---- pass 2: hand the records on with their new flag word (:443-465)
    uint64_t i = 0, written = 0;
    for (size_t b = 0; b < batches.size(); b++) {
        Batch * bt = batches[b];
        for (size_t k = 0; k + 1 < bt->offsets.size(); k++, i++) {
            const uint16_t flag = flags[i];
            if (removeDuplicates && (flag & 0x400)) continue;
            OGERead * al = rebuild_read(bt->data + bt->offsets[k], flag);
            putOutputAlignment(al);
            if (verbose && read_count && ++written % 100000 == 0)
                cerr << "\rWritten " << written << " records (" << written * 100 / read_count << "%)." << std::flush;
        }
        oge_gpu_host_free(bt->data);
        delete bt;
    }
    if (verbose && read_count) cerr << "\rWritten " << written << " records (" << written * 100 / read_count << "%)." << endl;
    return 0;
}
