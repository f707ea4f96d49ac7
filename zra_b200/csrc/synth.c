/* synth.c — deterministic synthetic inputs for tests and bench.py.
 *
 *  text-like : Zipf word model — 5000 words of length 2..10 drawn from a per-word random subset
 *              of the frequency-ordered alphabet "etaoinshrdlucmfwypvbgkqjxz", word i has weight
 *              1/(i+1), words joined by single spaces, '\n' after every 20000th word.
 *              The vocabulary depends only on vocabSeed; the word stream on streamSeed, so that
 *              independently generated chunks look like one corpus.
 *  random    : xoshiro256** bytes (incompressible).
 * Built into zra_b200/libzra_synth.so; no dependency on the CUDA library or on oracle/.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define SYNTH_API __attribute__((visibility("default")))
#define VOCAB 5000

static uint64_t splitmix(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
typedef struct { uint64_t s[4]; } xo;
static void xo_seed(xo* g, uint64_t seed) { for (int i = 0; i < 4; i++) g->s[i] = splitmix(&seed); }
static uint64_t xo_next(xo* g) {
    uint64_t r = rotl(g->s[1] * 5, 7) * 9, t = g->s[1] << 17;
    g->s[2] ^= g->s[0]; g->s[3] ^= g->s[1]; g->s[1] ^= g->s[2]; g->s[0] ^= g->s[3];
    g->s[2] ^= t; g->s[3] = rotl(g->s[3], 45);
    return r;
}

SYNTH_API void zra_synth_random(uint8_t* out, size_t n, uint64_t seed) {
    xo g; xo_seed(&g, seed);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t v = xo_next(&g); memcpy(out + i, &v, 8); }
    if (i < n) { uint64_t v = xo_next(&g); memcpy(out + i, &v, n - i); }
}

SYNTH_API void zra_synth_text(uint8_t* out, size_t n, uint64_t vocabSeed, uint64_t streamSeed) {
    static const char alphabet[] = "etaoinshrdlucmfwypvbgkqjxz";
    static char words[VOCAB][12];
    static uint8_t wlen[VOCAB];
    static uint32_t cdf[VOCAB]; /* cumulative weights scaled to 2^32 */
    xo g;
    /* vocabulary (rebuilt on every call: cheap, keeps the function re-entrant enough for tests) */
    char lw[VOCAB][12]; uint8_t ll[VOCAB]; uint32_t lc[VOCAB];
    xo_seed(&g, vocabSeed);
    for (int w = 0; w < VOCAB; w++) {
        int len = 2 + (int)(xo_next(&g) % 9);
        int subset = 4 + (int)(xo_next(&g) % 23); /* this word uses the `subset` most frequent letters */
        for (int k = 0; k < len; k++) {
            /* squared draw skews towards the frequent end of the subset */
            uint64_t r = xo_next(&g) % (uint64_t)(subset * subset);
            int idx = 0; while ((uint64_t)(idx + 1) * (idx + 1) <= r) idx++;
            lw[w][k] = alphabet[subset - 1 - idx < 0 ? 0 : subset - 1 - idx];
        }
        ll[w] = (uint8_t)len;
    }
    double total = 0; for (int w = 0; w < VOCAB; w++) total += 1.0 / (w + 1);
    double acc = 0;
    for (int w = 0; w < VOCAB; w++) { acc += 1.0 / (w + 1); double f = acc / total * 4294967296.0; lc[w] = f >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)f; }
    (void)words; (void)wlen; (void)cdf;
    xo_seed(&g, streamSeed ^ 0xA5A5A5A55A5A5A5AULL);
    size_t pos = 0; unsigned count = 0;
    while (pos < n) {
        uint32_t r = (uint32_t)(xo_next(&g) >> 32);
        int lo = 0, hi = VOCAB - 1;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (lc[mid] < r) lo = mid + 1; else hi = mid; }
        int len = ll[lo];
        for (int k = 0; k < len && pos < n; k++) out[pos++] = (uint8_t)lw[lo][k];
        if (pos < n) out[pos++] = (++count % 20000 == 0) ? '\n' : ' ';
    }
}
