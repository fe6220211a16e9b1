/*
 * apus_cert.h -- the checksum of a self-certifying publish, shared by the kernels
 * (apus_kernels.cu: leader_express computes it over the bytes it pushes, the follower over the bytes it finds) and by
 * the CPU property test (tests/hostlogic/cert_props.c), which is why it compiles as plain C as well.
 *
 * A lone request is pushed to the followers WITHOUT a writer-side fence (a system fence costs 1.5-1.7 us here, more than
 * the NVLink hop it orders): the publish record carries a checksum of the entry bytes, the follower re-reads the bytes
 * from its own HBM until they add up (FaRM-style object validation).  The checksum is LINEAR over the 8-byte words of
 * exactly the bytes [a, b) with position-dependent ODD weights (so a change confined to one word always shows: an odd
 * weight is invertible mod 2^64), and the record adds a key derived from the publish's count|term, so "the bytes that
 * were there before" only pass if they are the bytes that were sent -- in which case accepting them is harmless -- and a
 * certificate of an older publish does not verify against a newer one.
 */
#ifndef APUS_CERT_H
#define APUS_CERT_H
#include <stdint.h>
#ifdef __CUDACC__
#define APUS_HD __host__ __device__ __forceinline__
#else
#define APUS_HD static inline
#endif

APUS_HD uint64_t cs_weight(uint64_t word_index)
{
    uint64_t z = word_index * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    return z | 1ull;
}
/* bytes of the 8-byte word at log offset o that lie inside [a, b) */
APUS_HD uint64_t cs_mask(uint64_t o, uint64_t a, uint64_t b)
{
    const uint64_t lo = a > o ? a - o : 0, hi = b < o + 8 ? (b > o ? b - o : 0) : 8;
    if (hi <= lo) return 0;
    const uint64_t mh = hi >= 8 ? ~0ull : ((1ull << (8 * hi)) - 1ull);
    const uint64_t ml = (1ull << (8 * lo)) - 1ull;          /* lo < 8 here */
    return mh & ~ml;
}
/* contribution of the 16 B chunk {w0, w1} at log offset lo (16 B aligned), restricted to the bytes inside [a, b) */
APUS_HD uint64_t cs_chunk_words(uint64_t w0, uint64_t w1, uint64_t lo, uint64_t a, uint64_t b)
{
    return (w0 & cs_mask(lo, a, b)) * cs_weight(lo >> 3) + (w1 & cs_mask(lo + 8, a, b)) * cs_weight((lo >> 3) + 1);
}
/* the key that ties a certificate to ITS publish (a certificate half from an older publish must not verify) */
APUS_HD uint64_t cs_key(uint64_t cum_term) { return cs_weight(cum_term ^ 0x5851F42D4C957F2Dull); }
#endif
