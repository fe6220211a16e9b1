/*
 * tests/hostlogic/cert_props.c -- properties of the self-certifying publish's checksum (apus_b200/csrc/apus_cert.h, the
 * very functions the kernels use), checked on the CPU:
 *   1. chunk-wise (what the warps compute: one 16 B chunk per lane, masked to [a, b)) == a byte-wise definition
 *   2. a change of any single byte inside [a, b) changes the sum -- always (odd weights are invertible mod 2^64)
 *   3. bytes outside [a, b) -- the rest of the first and last chunk -- do not matter
 *   4. what the log held before (another entry's bytes at the same place) does not verify, unless it IS the same bytes
 *   5. the same bytes at another place do not verify (position-dependent weights)
 *   6. a certificate is tied to its publish: the key of another count or term does not verify
 * Prints "cert ok <cases>"; any violation aborts with a message.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../apus_b200/csrc/apus_cert.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

/* what a warp computes: the 16 B chunks that overlap [a, b), each restricted to the bytes inside */
static uint64_t sum_chunks(const uint8_t *log, uint64_t a, uint64_t b)
{
    uint64_t cs = 0;
    for (uint64_t lo = a & ~15ull; lo < b; lo += 16) {
        uint64_t w0, w1;
        memcpy(&w0, log + lo, 8); memcpy(&w1, log + lo + 8, 8);
        cs += cs_chunk_words(w0, w1, lo, a, b);
    }
    return cs;
}
/* the definition, byte by byte: byte at offset o contributes  byte << 8*(o % 8)  times the weight of word o / 8 */
static uint64_t sum_bytes(const uint8_t *log, uint64_t a, uint64_t b)
{
    uint64_t cs = 0;
    for (uint64_t o = a; o < b; o++) cs += ((uint64_t)log[o] << (8 * (o & 7))) * cs_weight(o >> 3);
    return cs;
}
#define CHECK(c, ...) do { if (!(c)) { fprintf(stderr, "FAILED line %d: ", __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } } while (0)

int main(void)
{
    enum { L = 1 << 16 };
    static uint8_t log[L + 32], old[L + 32], tmp[L + 32];
    unsigned long cases = 0;
    for (int round = 0; round < 4000; round++) {
        for (size_t i = 0; i < sizeof log; i++) { log[i] = (uint8_t)rnd(); old[i] = (uint8_t)rnd(); }
        const uint64_t len = 64 + rnd() % 700;                       /* an entry: header + a request's data image */
        const uint64_t a = (rnd() % (L - len - 64)) & ~(round & 1 ? 7ull : 0ull);   /* odd offsets too (ragged payloads) */
        const uint64_t b = a + len;
        const uint64_t cs = sum_chunks(log, a, b);
        CHECK(cs == sum_bytes(log, a, b), "chunk-wise != byte-wise for [%llu, %llu)", (unsigned long long)a, (unsigned long long)b);
        /* 2: every single-byte change inside shows (a few positions per entry, all 255 other values for one of them) */
        for (int k = 0; k < 12; k++) {
            const uint64_t o = a + rnd() % len;
            const uint8_t keep = log[o];
            const int all = k == 0;
            for (int d = 1; d < (all ? 256 : 2); d++) {
                log[o] = (uint8_t)(keep + (all ? d : 1 + rnd() % 255));
                CHECK(sum_chunks(log, a, b) != cs, "a changed byte at %llu went unnoticed", (unsigned long long)o);
                cases++;
            }
            log[o] = keep;
        }
        /* 3: the neighbours inside the first and last chunk do not matter */
        memcpy(tmp, log, sizeof log);
        for (uint64_t o = a & ~15ull; o < a; o++) tmp[o] ^= 0xFF;
        for (uint64_t o = b; o < ((b + 15) & ~15ull); o++) tmp[o] ^= 0xFF;
        CHECK(sum_chunks(tmp, a, b) == cs, "bytes outside [a, b) changed the sum");
        /* 4: what was there before does not verify; the same bytes do */
        CHECK(sum_chunks(old, a, b) != cs, "another entry's bytes verified");
        memcpy(old + a, log + a, len);
        CHECK(sum_chunks(old, a, b) == cs, "the same bytes did not verify");
        /* 5: the same bytes somewhere else (one word / one chunk / one entry further) */
        const uint64_t shifts[3] = { 8, 16, len };
        for (int s = 0; s < 3; s++) {
            const uint64_t a2 = a + shifts[s];
            if (a2 + len > L) continue;
            memset(tmp, 0, sizeof tmp);
            memcpy(tmp + a2, log + a, len);
            CHECK(sum_chunks(tmp, a2, a2 + len) != cs, "the same bytes at another offset verified");
        }
        /* 6: the key -- count | term << 48 */
        const uint64_t cum = rnd() % (1ull << 40), term = 1 + rnd() % 1000;
        const uint64_t rec = cs + cs_key(cum | term << 48);
        CHECK(sum_chunks(log, a, b) + cs_key(cum | term << 48) == rec, "self");
        CHECK(sum_chunks(log, a, b) + cs_key((cum + 1) | term << 48) != rec, "the certificate of the next publish verified");
        CHECK(sum_chunks(log, a, b) + cs_key(cum | (term + 1) << 48) != rec, "the certificate of another term verified");
        cases += 8;
    }
    /* weights are odd, and different for neighbouring words (what 2 and 5 rest on) */
    for (uint64_t w = 0; w < 200000; w++) { CHECK(cs_weight(w) & 1, "even weight"); CHECK(cs_weight(w) != cs_weight(w + 1) && cs_weight(w) != cs_weight(w + 2), "equal weights"); }
    printf("cert ok %lu\n", cases);
    return 0;
}
