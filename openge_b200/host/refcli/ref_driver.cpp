// Command-line front-end of the reference's own pipeline, for builds without Boost: it stands in for the argument parsing
// of commands/command_dedup.cpp and command_mergesort.cpp and nothing else.  Two users: oracle/ref_build/Makefile links it
// with the UNMODIFIED reference classes (-> oracle/_ref/oge_ref_dedup, the compiled reference: test infrastructure), and
// openge_b200/host/Makefile links it with this repo's drop-in MarkDuplicates / ReadSorter (-> oge_dedup_gpu,
// oge_mergesort_gpu).  It holds no duplicate-marking logic of its own.
//
// It wires the reference classes (compiled in place from /root/reference/openge/src, see the Makefiles) the way
// `openge dedup` does:
//   command_dedup.cpp:48-69   single chain  FileReader -> MarkDuplicates -> FileWriter
//   command_dedup.cpp:70-113  split chains  FileReader -> SplitByChromosome -> {MarkDuplicates} -> SortedMerge -> FileWriter
//   commands.cpp:59-84,110-112  verbose / threads / pool set-up and tear-down
// Boost is absent here, so the reference's own CLI front-end cannot be built; this
// file replaces only that argument parsing.
//
// Modes
//   file (default):  in.bam -> out file (bam or rawbam by extension / -F)
//   --mem:           preload all records into RAM, then time
//                    MemorySource -> MarkDuplicates -> CountingSink   (records in, flags out)
//                    i.e. exactly MarkDuplicates::runInternal with host buffers on both sides.
//                    Prints one JSON line with seconds per repetition.
//   --sort:          (single chain only) `openge mergesort [-M]` (commands/command_mergesort.cpp:70-113): the reference's
//                    ReadSorter (algorithms/read_sorter.cpp, coordinate order) in front; --nodedup leaves MarkDuplicates out;
//                    -n N = alignments per temp file (default 200000 as there)
//   --stats:         (single chain only) puts the reference's Statistics module (algorithms/statistics.cpp,
//                    what `openge stats` runs, command_stats.cpp) between MarkDuplicates and the writer:
//                    its report goes to stdout.
#include "algorithms/algorithm_module.h"
#include "algorithms/file_reader.h"
#include "algorithms/file_writer.h"
#include "algorithms/mark_duplicates.h"
#include "algorithms/sorted_merge.h"
#include "algorithms/split_by_chromosome.h"
#include "algorithms/statistics.h"
#include "algorithms/read_sorter.h"
#include "util/read_stream_reader.h"

#include <sys/time.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

using namespace std;

static double now_s() {
    timeval t; gettimeofday(&t, NULL);
    return t.tv_sec + 1e-6 * t.tv_usec;
}

// Feeds preloaded records (copies, because the receiving module deletes them).
class MemorySource : public AlgorithmModule {
public:
    vector<OGERead *> * reads;
    BamHeader header;
    virtual const BamHeader & getHeader() { return header; }
protected:
    virtual int runInternal() {
        for (size_t i = 0; i < reads->size(); i++) {
            OGERead * al = OGERead::allocate();
            *al = *(*reads)[i];
            putOutputAlignment(al);
        }
        return 0;
    }
};

// Terminal stage: records the flag word of every record it receives.
class CountingSink : public AlgorithmModule {
public:
    vector<uint16_t> flags;
    size_t dups;
    CountingSink() : dups(0) {}
protected:
    virtual int runInternal() {
        while (true) {
            OGERead * r = getInputAlignment();
            if (!r) break;
            flags.push_back((uint16_t) r->getAlignmentFlag());
            if (r->IsDuplicate()) dups++;
            putOutputAlignment(r);
        }
        return 0;
    }
};

static void usage() {
    fprintf(stderr,
        "usage: oge_ref_dedup [-v] [--nosplit] [-r] [-t N] [-T tmpdir] [-F fmt] [-c lvl] [--stats] [--mem [--reps K] [--flags out.u16]] in out\n");
    exit(2);
}

int main(int argc, char ** argv) {
    bool verbose = false, nosplit = false, remove_dups = false, mem = false, stats = false, sort = false, nodedup = false;
    int per_tempfile = 200000;
    int threads = ThreadPool::availableCores();
    int level = 6, reps = 1;
    string tmpdir = "/tmp", format, flags_out;
    vector<string> pos;
    for (int i = 1; i < argc; i++) {
        string a = argv[i];
        if (a == "-v") verbose = true;
        else if (a == "--nosplit") nosplit = true;
        else if (a == "-r") remove_dups = true;
        else if (a == "--mem") mem = true;
        else if (a == "--stats") stats = true;
        else if (a == "--sort") sort = true;
        else if (a == "--nodedup") nodedup = true;
        else if (a == "-n" && i + 1 < argc) per_tempfile = atoi(argv[++i]);
        else if (a == "-t" && i + 1 < argc) threads = atoi(argv[++i]);
        else if (a == "-T" && i + 1 < argc) tmpdir = argv[++i];
        else if (a == "-F" && i + 1 < argc) format = argv[++i];
        else if (a == "-c" && i + 1 < argc) level = atoi(argv[++i]);
        else if (a == "--reps" && i + 1 < argc) reps = atoi(argv[++i]);
        else if (a == "--flags" && i + 1 < argc) flags_out = argv[++i];
        else if (a[0] == '-') usage();
        else pos.push_back(a);
    }
    if (pos.size() < (mem ? 1u : 2u)) usage();
    tmpdir += "/";

    // commands.cpp:59-84
    OGEParallelismSettings::setNumberThreads(threads);
    AlgorithmModule::setNothreads(false);
    AlgorithmModule::setVerbose(verbose);
    OGEParallelismSettings::enableMultithreading();

    int num_chains = min(12, OGEParallelismSettings::getNumberThreads() / 2);   // command_dedup.cpp:46
    int ret = 0;

    if (mem) {
        vector<OGERead *> reads;
        MultiReader reader;
        if (!reader.open(pos[0])) { fprintf(stderr, "cannot open %s\n", pos[0].c_str()); return 1; }
        BamHeader header = reader.getHeader();
        while (true) { OGERead * r = reader.read(); if (!r) break; reads.push_back(r); }
        reader.close();

        printf("{\"records\": %zu, \"threads\": %d, \"seconds\": [", reads.size(), threads);
        size_t dups = 0;
        for (int rep = 0; rep < reps; rep++) {
            MemorySource src; src.reads = &reads; src.header = header;
            MarkDuplicates md(tmpdir);
            CountingSink sink;
            sink.flags.reserve(reads.size());
            src.addSink(&md);
            md.addSink(&sink);
            md.removeDuplicates = remove_dups;
            double t0 = now_s();
            sink.runChain();
            double t1 = now_s();
            dups = sink.dups;
            printf("%s%.6f", rep ? ", " : "", t1 - t0);
            if (rep == reps - 1 && !flags_out.empty()) {
                FILE * f = fopen(flags_out.c_str(), "wb");
                fwrite(&sink.flags[0], 2, sink.flags.size(), f);
                fclose(f);
            }
        }
        printf("], \"duplicates\": %zu}\n", dups);
    } else if (sort) {
        FileReader reader;
        ReadSorter sort_reads(tmpdir);
        MarkDuplicates mark_duplicates(tmpdir);
        FileWriter writer;
        reader.addSink(&sort_reads);
        if (!nodedup) {
            sort_reads.addSink(&mark_duplicates);
            mark_duplicates.addSink(&writer);
            mark_duplicates.removeDuplicates = remove_dups;
        } else {
            sort_reads.addSink(&writer);
        }
        sort_reads.setSortBy(BamHeader::SORT_COORDINATE);
        sort_reads.setCompressTempFiles(false);
        sort_reads.setAlignmentsPerTempfile(per_tempfile);
        if (!format.empty()) writer.setFormat(format);
        reader.addFile(pos[0]);
        writer.setFilename(pos[1]);
        writer.setCompressionLevel(level);
        ret = writer.runChain();
    } else if (nosplit || num_chains <= 1) {
        FileReader reader;
        MarkDuplicates mark_duplicates(tmpdir);
        FileWriter writer;
        Statistics statistics;
        reader.addSink(&mark_duplicates);
        if (!format.empty()) writer.setFormat(format);
        if (stats) {
            statistics.showReadLengthSummary(false);      // the member is left uninitialised by the constructor (statistics.cpp:32-48)
            mark_duplicates.addSink(&statistics);
            statistics.addSink(&writer);
        } else {
            mark_duplicates.addSink(&writer);
        }
        mark_duplicates.removeDuplicates = remove_dups;
        reader.addFile(pos[0]);
        writer.setFilename(pos[1]);
        writer.setCompressionLevel(level);
        ret = writer.runChain();
    } else {
        FileReader reader;
        SortedMerge merge;
        SplitByChromosome split;
        FileWriter writer;
        vector<MarkDuplicates *> markers;
        reader.addSink(&split);
        merge.addSink(&writer);
        for (int c = 0; c < num_chains; c++) {
            MarkDuplicates * md = new MarkDuplicates(tmpdir);
            markers.push_back(md);
            merge.addSource(md);
            md->removeDuplicates = remove_dups;
            split.addSink(md);
        }
        if (!format.empty()) writer.setFormat(format);
        reader.addFile(pos[0]);
        writer.setFilename(pos[1]);
        writer.setCompressionLevel(level);
        ret = writer.runChain();
        for (size_t c = 0; c < markers.size(); c++) delete markers[c];
    }

    // commands.cpp:110-112
    OGERead::clearCachedAllocations();
    ThreadPool::closeSharedPool();
    return ret;
}
