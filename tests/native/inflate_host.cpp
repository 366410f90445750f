// TEST HARNESS: compiles the product's DEFLATE decoder (openge_b200/csrc/inflate_core.cuh) for the host with one
// lane, so that tests/test_inflate_core.py can run it against zlib without a GPU.
#include <stdlib.h>
#include <string.h>

#include "inflate_core.cuh"

extern "C" int oge_test_inflate_block(const unsigned char *in, unsigned in_len, unsigned char *out, unsigned out_len) {
    oge_inflate::Tables *T = (oge_inflate::Tables *) calloc(1, sizeof(oge_inflate::Tables));
    // the decoder may read up to 12 bytes past the payload (in a BGZF file the footer and the next header are there)
    unsigned char *padded = (unsigned char *) calloc(1, (size_t) in_len + 32);
    memcpy(padded, in, in_len);
    const int rc = oge_inflate::inflate_block(padded, in_len, out, out_len, T, 0);
    free(padded);
    free(T);
    return rc;
}

extern "C" unsigned oge_test_inflate_tables_bytes(void) { return (unsigned) sizeof(oge_inflate::Tables); }
