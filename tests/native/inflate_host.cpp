// TEST HARNESS: compiles the product's DEFLATE decoder (openge_b200/csrc/inflate_core.cuh) for the host with one
// lane, so that tests/test_inflate_core.py can run it against zlib without a GPU.
#include <stdlib.h>
#include <string.h>

#include "inflate_core.cuh"

extern "C" int oge_test_inflate_block2(const unsigned char *in, unsigned in_len, unsigned char *out, unsigned out_len, int small_tables) {
    oge_inflate::Tables *T = (oge_inflate::Tables *) calloc(1, sizeof(oge_inflate::Tables));
    // the decoder may read up to 12 bytes past the payload (in a BGZF file the footer and the next header are there)
    unsigned char *padded = (unsigned char *) calloc(1, (size_t) in_len + 32);
    memcpy(padded, in, in_len);
    int rc;
    if (small_tables == 2)      // the state-machine form of the thread-per-block kernel, with its table widths
        rc = oge_inflate::inflate_lockstep<9, 7>(padded, in_len, out, out_len, oge_inflate::tables_ref(T), true);
    else if (small_tables)
        rc = oge_inflate::inflate_block<1, 9, 7>(padded, in_len, out, out_len, oge_inflate::tables_ref(T), 0);
    else
        rc = oge_inflate::inflate_block<1, oge_inflate::LIT_BITS, oge_inflate::DIST_BITS>(padded, in_len, out, out_len, oge_inflate::tables_ref(T), 0);
    free(padded);
    free(T);
    return rc;
}

extern "C" int oge_test_inflate_block(const unsigned char *in, unsigned in_len, unsigned char *out, unsigned out_len) {
    return oge_test_inflate_block2(in, in_len, out, out_len, 0);
}

extern "C" unsigned oge_test_inflate_tables_bytes(void) { return (unsigned) sizeof(oge_inflate::Tables); }
