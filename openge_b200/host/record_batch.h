// Shared by the drop-in classes (mark_duplicates_gpu.cpp, read_sorter_gpu.cpp): reads of the reference's pipeline framed
// into pinned host batches in the raw BAM record layout the C ABI takes (util/bam_serializer.h:106-141: block_size, 32-byte
// core, name, cigar, packed bases, qualities, tags), and rebuilt from it.  gnu++98 like the rest of the reference.
#ifndef OGE_HOST_RECORD_BATCH_H
#define OGE_HOST_RECORD_BATCH_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <iostream>
#include <string>
#include <vector>

#include "oge_gpu_dedup.h"
#include "util/oge_read.h"

namespace oge_host {

const size_t BATCH_BYTES = (size_t) 128 << 20;      // one pinned staging buffer

struct Batch {
    uint8_t * data;                  // pinned (oge_gpu_host_alloc)
    size_t used;
    std::vector<uint64_t> offsets;   // n + 1, relative to data
    Batch() : data(NULL), used(0) { offsets.push_back(0); }
};

inline void put_u32(uint8_t * p, uint32_t v) { memcpy(p, &v, 4); }
inline uint32_t get_u32(const uint8_t * p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t get_i32(const uint8_t * p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t get_u16(const uint8_t * p) { uint16_t v; memcpy(&v, p, 2); return v; }

// bytes the read takes in a batch
inline size_t record_bytes(const OGERead & al) { return 4 + 32 + al.getSupportData().getAllCharData().size(); }

// the read appended to the batch (the caller has made room)
inline void append_read(Batch & b, const OGERead & al) {
    const std::string & chars = al.getSupportData().getAllCharData();
    uint8_t * p = b.data + b.used;
    put_u32(p, (uint32_t) (32 + chars.size()));
    put_u32(p + 4, (uint32_t) al.getRefID());
    put_u32(p + 8, (uint32_t) al.getPosition());
    put_u32(p + 12, ((uint32_t) al.getBin() << 16) | ((uint32_t) (al.getMapQuality() & 0xFF) << 8) | (uint32_t) (al.getNameLength() & 0xFF));
    put_u32(p + 16, ((uint32_t) al.getAlignmentFlag() << 16) | (uint32_t) (al.getNumCigarOps() & 0xFFFF));
    put_u32(p + 20, (uint32_t) al.getLength());
    put_u32(p + 24, (uint32_t) al.getMateRefID());
    put_u32(p + 28, (uint32_t) al.getMatePosition());
    put_u32(p + 32, (uint32_t) al.getInsertSize());
    memcpy(p + 36, chars.data(), chars.size());
    b.used += 4 + 32 + chars.size();
    b.offsets.push_back(b.used);
}

// a read of the pipeline from one framed record, with the given flag word
inline OGERead * rebuild_read(const uint8_t * p, uint16_t flag) {
    const uint32_t block = get_u32(p);
    OGERead * al = OGERead::allocate();
    al->setRefID(get_i32(p + 4));
    al->setPosition(get_i32(p + 8));
    al->setMapQuality(p[13]);
    al->setBin(get_u16(p + 14));
    al->setAlignmentFlag(flag);
    al->setMateRefID(get_i32(p + 24));
    al->setMatePosition(get_i32(p + 28));
    al->setInsertSize(get_i32(p + 32));
    al->setBamStringData((const char *) p + 36, block - 32, get_u16(p + 16), get_u32(p + 20), p[12]);
    return al;
}

}  // namespace oge_host
#endif
