/*
 * zra_oracle.c — CPU restatement of the ZRA hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the checker for the CUDA implementation in zra_b200/csrc. It is
 * a plain, serial, deliberately unoptimised C restatement of what the reference
 * computes on this path; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it. The product
 * library (libzra_b200.so) never links, loads or calls anything in oracle/.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function here against
 *   - the reference itself, compiled from /root/reference by oracle/Makefile
 *     into oracle/_ref/libzra_ref.so (zra::CompressBuffer/DecompressBuffer/
 *     DecompressRA, ZSTD_decompress, XXH64, CRC),
 *   - the committed fixtures under tests/golden/ (reference-made archives,
 *     decodecorpus frames, zstd's golden-decompression file),
 *   - known answers (CRC-32 check value 0xCBF43926, XXH64 vectors).
 *
 * What is restated, with the reference location each part follows
 * (paths relative to /root/reference, zstd/ = submodule/zstd/):
 *   CRC-32 ................ submodule/CRCpp/inc/CRC.h:434-462 (bitwise), params :1565-1569
 *   XXH64 ................. zstd/lib/common/xxhash.c:415-500, 567-720
 *   backward bitstream .... zstd/lib/common/bitstream.h:272-440
 *   FSE NCount reader ..... zstd/lib/common/entropy_common.c:41-145
 *   FSE decode table ...... zstd/lib/decompress/zstd_decompress_block.c:367-427,
 *                           zstd/lib/common/fse_decompress.c:66-130
 *   Huffman weights ....... zstd/lib/common/entropy_common.c:155-216,
 *                           zstd/lib/common/fse_decompress.c:177-273
 *   Huffman decode ........ zstd/lib/decompress/huf_decompress.c:118-354
 *   literals section ...... zstd/lib/decompress/zstd_decompress_block.c:79-235
 *   sequence headers ...... zstd/lib/decompress/zstd_decompress_block.c:433-550
 *   sequence decode ....... zstd/lib/decompress/zstd_decompress_block.c:795-948
 *   sequence execute ...... zstd/lib/decompress/zstd_decompress_block.c:576-793, 999-1117
 *   frame / multi-frame ... zstd/lib/decompress/zstd_decompress.c:244-318, 609-785
 *   ZRA header / table .... source/zra.cpp:96-171
 *   ZRA sizes ............. source/zra.cpp:189-198, zstd/lib/zstd.h:174
 *   ZRA DecompressBuffer .. source/zra.cpp:243-250
 *   ZRA DecompressRA ...... source/zra.cpp:258-296 (and the streaming twin :369-413)
 *
 * Return convention: every decoding function returns a non-negative byte count
 * or the NEGATED ZSTD_ErrorCode (zstd/lib/common/zstd_errors.h:52-79).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

enum {
    E_GENERIC = 1, E_PREFIX_UNKNOWN = 10, E_FRAMEPARAM_UNSUPPORTED = 14, E_WINDOW_TOO_LARGE = 16,
    E_CORRUPTION = 20, E_CHECKSUM_WRONG = 22, E_DICT_CORRUPTED = 30, E_DICT_WRONG = 32,
    E_TABLELOG_TOO_LARGE = 44, E_MAXSYM_TOO_SMALL = 48, E_DST_TOO_SMALL = 70, E_SRC_WRONG = 72
};

#define BLOCKSIZE_MAX (1u << 17)
#define LONGNBSEQ 0x7F00

typedef int64_t i64;
typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

static u32 rd16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
static u32 rd24(const u8* p) { return rd16(p) | ((u32)p[2] << 16); }
static u32 rd32(const u8* p) { return rd16(p) | (rd16(p + 2) << 16); }
static u64 rd64(const u8* p) { return (u64)rd32(p) | ((u64)rd32(p + 4) << 32); }
static int highbit(u32 v) { int r = -1; while (v) { v >>= 1; r++; } return r; }

/* ------------------------------------------------------------------ CRC-32 */
/* Bitwise, reflected, poly 0x04C11DB7 (0xEDB88320 reflected), init/xorout all-ones.
 * `prev` is the finished CRC of the preceding bytes (0 for none), as CRCpp chains it. */
ORACLE_API u32 zra_oracle_crc32(const void* data, size_t n, u32 prev) {
    const u8* p = (const u8*)data;
    u32 c = prev ^ 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
    }
    return c ^ 0xFFFFFFFFu;
}

/* ------------------------------------------------------------------- XXH64 */
#define P1 11400714785074694791ULL
#define P2 14029467366897019727ULL
#define P3 1609587929392839161ULL
#define P4 9650029242287828579ULL
#define P5 2870177450012600261ULL
static u64 rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
static u64 xxh_round(u64 acc, u64 in) { return rotl64(acc + in * P2, 31) * P1; }
static u64 xxh_merge(u64 acc, u64 v) { return (acc ^ xxh_round(0, v)) * P1 + P4; }

ORACLE_API u64 zra_oracle_xxh64(const void* data, size_t n, u64 seed) {
    const u8* p = (const u8*)data;
    const u8* end = p + n;
    u64 h;
    if (n >= 32) {
        u64 v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do {
            v1 = xxh_round(v1, rd64(p));
            v2 = xxh_round(v2, rd64(p + 8));
            v3 = xxh_round(v3, rd64(p + 16));
            v4 = xxh_round(v4, rd64(p + 24));
            p += 32;
        } while (p + 32 <= end);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
    } else {
        h = seed + P5;
    }
    h += (u64)n;
    while (p + 8 <= end) { h ^= xxh_round(0, rd64(p)); h = rotl64(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (u64)rd32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (*p++) * P5; h = rotl64(h, 11) * P1; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* ------------------------------------------------------ backward bitstream */
/* `bits` = number of not-yet-read bits, numbered LSB-first from the start of the
 * buffer. Reading past the start yields zero bits and drives `bits` negative. */
typedef struct { const u8* p; i64 bits; } bstream;

static int bs_init(bstream* b, const u8* p, size_t len) {
    if (len < 1) return -E_SRC_WRONG;
    if (p[len - 1] == 0) return -E_CORRUPTION; /* missing end mark */
    b->p = p;
    b->bits = (i64)(len - 1) * 8 + highbit(p[len - 1]);
    return 0;
}
static u64 bs_peek(const bstream* b, int n) {
    /* the n bits just below the read cursor, most significant first */
    i64 lo = b->bits - n;
    u64 v = 0;
    for (int i = n - 1; i >= 0; i--) {
        i64 pos = lo + i;
        u64 bit = (pos < 0) ? 0 : ((b->p[pos >> 3] >> (pos & 7)) & 1u);
        v = (v << 1) | bit;
    }
    return v;
}
static u64 bs_read(bstream* b, int n) { u64 v = bs_peek(b, n); b->bits -= n; return v; }

/* --------------------------------------------------------------------- FSE */
typedef struct { u16 nextState; u8 nbBits; u8 symbol; u8 nbAddBits; u32 baseValue; } fse_entry;
typedef struct { fse_entry e[512]; int log; } fse_table;

/* Reads a normalised-count header. Returns bytes consumed or a negative error. */
static int fse_read_ncount(int16_t* norm, unsigned* maxSymbol, unsigned* tableLog, const u8* src, size_t len) {
    /* forward LSB-first bit reader over the header bytes */
    u64 pos = 0;
    const u64 limit = (u64)len * 8;
#define NC_PEEK(n, out) do { u32 _v = 0; for (int _i = 0; _i < (n); _i++) { u64 _q = pos + _i; \
        u32 _b = (_q < limit) ? ((src[_q >> 3] >> (_q & 7)) & 1u) : 0; _v |= _b << _i; } (out) = _v; } while (0)
    if (len < 1) return -E_SRC_WRONG;
    u32 v;
    NC_PEEK(4, v); pos += 4;
    unsigned log = v + 5;
    if (log > 15) return -E_TABLELOG_TOO_LARGE;
    *tableLog = log;
    int remaining = (1 << log) + 1;
    int threshold = 1 << log;
    int nbBits = (int)log + 1;
    unsigned sym = 0;
    int previous0 = 0;
    while (remaining > 1 && sym <= *maxSymbol) {
        if (previous0) {
            unsigned n0 = sym;
            for (;;) {
                NC_PEEK(2, v); pos += 2;
                n0 += v;
                if (v != 3) break;
            }
            if (n0 > *maxSymbol) return -E_MAXSYM_TOO_SMALL;
            while (sym < n0) norm[sym++] = 0;
        }
        {
            int max = (2 * threshold - 1) - remaining;
            int count;
            NC_PEEK(nbBits, v);
            if ((int)(v & (u32)(threshold - 1)) < max) {
                count = (int)(v & (u32)(threshold - 1));
                pos += (u64)(nbBits - 1);
            } else {
                count = (int)(v & (u32)(2 * threshold - 1));
                if (count >= threshold) count -= max;
                pos += (u64)nbBits;
            }
            count--;
            remaining -= count < 0 ? -count : count;
            norm[sym++] = (int16_t)count;
            previous0 = !count;
            while (remaining < threshold) { nbBits--; threshold >>= 1; }
        }
    }
#undef NC_PEEK
    if (remaining != 1) return -E_CORRUPTION;
    if (pos > limit) return -E_CORRUPTION;
    *maxSymbol = sym - 1;
    return (int)((pos + 7) >> 3);
}

/* Spreads symbols and derives (nbBits, nextState) for every state. */
static void fse_build(fse_table* t, const int16_t* norm, unsigned maxSymbol, unsigned log,
                      const u32* base, const u32* addBits) {
    u32 size = 1u << log, mask = size - 1, high = size - 1;
    u16 next[256];
    t->log = (int)log;
    for (unsigned s = 0; s <= maxSymbol; s++) {
        if (norm[s] == -1) { t->e[high--].symbol = (u8)s; next[s] = 1; }
        else next[s] = (u16)norm[s];
    }
    u32 step = (size >> 1) + (size >> 3) + 3, pos = 0;
    for (unsigned s = 0; s <= maxSymbol; s++) {
        for (int i = 0; i < norm[s]; i++) {
            t->e[pos].symbol = (u8)s;
            do { pos = (pos + step) & mask; } while (pos > high);
        }
    }
    for (u32 u = 0; u < size; u++) {
        u8 s = t->e[u].symbol;
        u32 ns = next[s]++;
        t->e[u].nbBits = (u8)(log - (unsigned)highbit(ns));
        t->e[u].nextState = (u16)((ns << t->e[u].nbBits) - size);
        t->e[u].baseValue = base ? base[s] : 0;
        t->e[u].nbAddBits = addBits ? (u8)addBits[s] : 0;
    }
}

static void fse_build_rle(fse_table* t, u8 sym, const u32* base, const u32* addBits) {
    t->log = 0;
    t->e[0].symbol = sym; t->e[0].nbBits = 0; t->e[0].nextState = 0;
    t->e[0].baseValue = base[sym]; t->e[0].nbAddBits = (u8)addBits[sym];
}

/* ----------------------------------------------------------------- Huffman */
typedef struct { u8 symbol; u8 nbBits; } huf_entry;
typedef struct { huf_entry e[4096]; int log; int valid; } huf_table;

/* Huffman weights compressed with FSE: two interleaved states, decode until the stream runs dry. */
static int huf_fse_weights(u8* w, int cap, const u8* src, size_t len) {
    int16_t norm[256];
    unsigned maxSym = 255, log;
    int h = fse_read_ncount(norm, &maxSym, &log, src, len);
    if (h < 0) return h;
    if (log > 6) return -E_TABLELOG_TOO_LARGE;
    fse_table t;
    fse_build(&t, norm, maxSym, log, NULL, NULL);
    bstream b;
    int r = bs_init(&b, src + h, len - (size_t)h);
    if (r < 0) return r;
    u32 s1 = (u32)bs_read(&b, (int)log), s2 = (u32)bs_read(&b, (int)log);
    int n = 0;
    for (;;) {
        if (n > cap - 2) return -E_DST_TOO_SMALL;
        w[n++] = t.e[s1].symbol;
        s1 = t.e[s1].nextState + (u32)bs_read(&b, t.e[s1].nbBits);
        if (b.bits < 0) { w[n++] = t.e[s2].symbol; break; }
        if (n > cap - 2) return -E_DST_TOO_SMALL;
        w[n++] = t.e[s2].symbol;
        s2 = t.e[s2].nextState + (u32)bs_read(&b, t.e[s2].nbBits);
        if (b.bits < 0) { w[n++] = t.e[s1].symbol; break; }
    }
    return n;
}

/* Parses a Huffman tree description and fills the single-symbol decode table.
 * Returns bytes consumed. */
static int huf_read_table(huf_table* t, const u8* src, size_t len) {
    u8 w[256];
    u32 rank[16] = {0};
    int n, consumed;
    if (len < 1) return -E_SRC_WRONG;
    int h = src[0];
    if (h >= 128) {
        n = h - 127;
        consumed = (n + 1) / 2;
        if ((size_t)consumed + 1 > len) return -E_SRC_WRONG;
        for (int i = 0; i < n; i += 2) { w[i] = src[1 + i / 2] >> 4; w[i + 1] = src[1 + i / 2] & 15; }
    } else {
        consumed = h;
        if ((size_t)consumed + 1 > len) return -E_SRC_WRONG;
        n = huf_fse_weights(w, 255, src + 1, (size_t)h);
        if (n < 0) return n;
    }
    u32 total = 0;
    for (int i = 0; i < n; i++) {
        if (w[i] >= 12) return -E_CORRUPTION;
        rank[w[i]]++;
        total += (1u << w[i]) >> 1;
    }
    if (total == 0) return -E_CORRUPTION;
    int log = highbit(total) + 1;
    if (log > 12) return -E_CORRUPTION;
    u32 rest = (1u << log) - total;
    int hb = highbit(rest);
    if ((1u << hb) != rest) return -E_CORRUPTION; /* last weight must be a clean power of two */
    w[n] = (u8)(hb + 1);
    rank[w[n]]++;
    n++;
    if (rank[1] < 2 || (rank[1] & 1)) return -E_CORRUPTION;
    /* canonical fill: ascending weight, then ascending symbol */
    u32 start[16], nxt = 0;
    for (int r = 1; r <= log; r++) { start[r] = nxt; nxt += rank[r] << (r - 1); }
    for (int s = 0; s < n; s++) {
        int ww = w[s];
        if (!ww) continue;
        u32 span = (1u << ww) >> 1;
        for (u32 u = start[ww]; u < start[ww] + span; u++) { t->e[u].symbol = (u8)s; t->e[u].nbBits = (u8)(log + 1 - ww); }
        start[ww] += span;
    }
    t->log = log;
    t->valid = 1;
    return consumed + 1;
}

static int huf_stream(const huf_table* t, u8* dst, size_t n, const u8* src, size_t len) {
    bstream b;
    int r = bs_init(&b, src, len);
    if (r < 0) return -E_CORRUPTION;
    for (size_t i = 0; i < n; i++) {
        huf_entry e = t->e[bs_peek(&b, t->log)];
        b.bits -= e.nbBits;
        dst[i] = e.symbol;
    }
    return b.bits == 0 ? 0 : -E_CORRUPTION;
}

static int huf_decode(const huf_table* t, u8* dst, size_t n, const u8* src, size_t len, int single) {
    if (single) return huf_stream(t, dst, n, src, len);
    if (len < 10) return -E_CORRUPTION;
    size_t l1 = rd16(src), l2 = rd16(src + 2), l3 = rd16(src + 4);
    if (l1 + l2 + l3 + 6 > len) return -E_CORRUPTION;
    size_t l4 = len - (l1 + l2 + l3 + 6);
    size_t seg = (n + 3) / 4;
    if (seg * 3 > n) return -E_CORRUPTION;
    const u8* s = src + 6;
    int r;
    if ((r = huf_stream(t, dst, seg, s, l1)) < 0) return r;
    if ((r = huf_stream(t, dst + seg, seg, s + l1, l2)) < 0) return r;
    if ((r = huf_stream(t, dst + 2 * seg, seg, s + l1 + l2, l3)) < 0) return r;
    return huf_stream(t, dst + 3 * seg, n - 3 * seg, s + l1 + l2 + l3, l4);
}

/* ---------------------------------------------------- sequence code tables */
static const u32 LL_base[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40,
                                48, 64, 0x80, 0x100, 0x200, 0x400, 0x800, 0x1000, 0x2000, 0x4000, 0x8000, 0x10000};
static const u32 LL_bits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3,
                                4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const u32 ML_base[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26,
                                27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 0x83, 0x103,
                                0x203, 0x403, 0x803, 0x1003, 0x2003, 0x4003, 0x8003, 0x10003};
static const u32 ML_bits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static u32 OF_base[32], OF_bits[32];
static const int16_t LL_defnorm[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2,
                                       2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const int16_t ML_defnorm[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                       1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static const int16_t OF_defnorm[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

static void of_init(void) {
    if (OF_bits[5]) return;
    for (u32 c = 0; c < 32; c++) { OF_bits[c] = c; OF_base[c] = (c < 2) ? c : ((1u << c) - 3); }
    OF_base[1] = 1;
}

/* ------------------------------------------------------------ frame decode */
typedef struct {
    huf_table huf;
    fse_table ll, of, ml;
    int fseValid;
    u32 rep[3];
    u8* lit; /* BLOCKSIZE_MAX */
} frame_ctx;

static i64 build_seq_table(fse_table* t, int type, unsigned maxSym, unsigned maxLog, const u8* src, size_t len,
                           const u32* base, const u32* bits, const int16_t* defnorm, unsigned defLog, int repeatOk) {
    switch (type) {
        case 1: /* RLE */
            if (!len) return -E_CORRUPTION;
            if (src[0] > maxSym) return -E_CORRUPTION;
            fse_build_rle(t, src[0], base, bits);
            return 1;
        case 0: /* predefined */
            fse_build(t, defnorm, maxSym == 35 ? 35 : (maxSym == 52 ? 52 : 28), defLog, base, bits);
            return 0;
        case 3: /* repeat */
            if (!repeatOk) return -E_CORRUPTION;
            return 0;
        default: {
            int16_t norm[64];
            unsigned log, ms = maxSym;
            int h = fse_read_ncount(norm, &ms, &log, src, len);
            if (h < 0) return -E_CORRUPTION;
            if (log > maxLog) return -E_CORRUPTION;
            fse_build(t, norm, ms, log, base, bits);
            return h;
        }
    }
}

/* One compressed block. `frameStart` is the first byte of this frame's output (match window floor). */
static i64 decode_block(frame_ctx* c, u8* dst, size_t cap, const u8* frameStart, const u8* src, size_t len) {
    if (len >= BLOCKSIZE_MAX) return -E_SRC_WRONG;
    if (len < 3) return -E_CORRUPTION;
    /* literals section */
    const u8* litPtr;
    size_t litSize, used;
    {
        int type = src[0] & 3, fmt = (src[0] >> 2) & 3;
        if (type >= 2) {
            size_t lh, cs;
            int single = 0;
            if (type == 3 && !c->huf.valid) return -E_DICT_CORRUPTED;
            if (len < 5) return -E_CORRUPTION;
            u32 w = rd32(src);
            if (fmt <= 1) { single = !fmt; lh = 3; litSize = (w >> 4) & 0x3FF; cs = (w >> 14) & 0x3FF; }
            else if (fmt == 2) { lh = 4; litSize = (w >> 4) & 0x3FFF; cs = w >> 18; }
            else { lh = 5; litSize = (w >> 4) & 0x3FFFF; cs = (w >> 22) + ((size_t)src[4] << 10); }
            if (litSize > BLOCKSIZE_MAX) return -E_CORRUPTION;
            if (cs + lh > len) return -E_CORRUPTION;
            const u8* hs = src + lh;
            size_t hl = cs;
            if (type == 2) {
                huf_table nt;
                nt.valid = 0;
                int h = huf_read_table(&nt, hs, hl);
                if (h < 0) return -E_CORRUPTION;
                c->huf = nt;
                hs += h; hl -= (size_t)h;
            }
            if (huf_decode(&c->huf, c->lit, litSize, hs, hl, single) < 0) return -E_CORRUPTION;
            litPtr = c->lit;
            used = lh + cs;
        } else {
            size_t lh;
            if (fmt == 0 || fmt == 2) { lh = 1; litSize = src[0] >> 3; }
            else if (fmt == 1) { lh = 2; litSize = rd16(src) >> 4; }
            else { lh = 3; litSize = rd24(src) >> 4; }
            if (type == 0) {
                if (lh + litSize > len) return -E_CORRUPTION;
                litPtr = src + lh;
                used = lh + litSize;
            } else {
                if (fmt == 3 && len < 4) return -E_CORRUPTION;
                if (litSize > BLOCKSIZE_MAX) return -E_CORRUPTION;
                memset(c->lit, src[lh], litSize);
                litPtr = c->lit;
                used = lh + 1;
            }
        }
    }
    src += used; len -= used;
    /* sequences header */
    if (len < 1) return -E_SRC_WRONG;
    const u8* ip = src;
    const u8* iend = src + len;
    int nbSeq = *ip++;
    if (!nbSeq) {
        if (len != 1) return -E_SRC_WRONG;
    } else {
        if (nbSeq > 0x7F) {
            if (nbSeq == 0xFF) { if (ip + 2 > iend) return -E_SRC_WRONG; nbSeq = (int)rd16(ip) + LONGNBSEQ; ip += 2; }
            else { if (ip >= iend) return -E_SRC_WRONG; nbSeq = ((nbSeq - 0x80) << 8) + *ip++; }
        }
        if (ip + 1 > iend) return -E_SRC_WRONG;
        int modes = *ip++;
        i64 h;
        of_init();
        h = build_seq_table(&c->ll, modes >> 6, 35, 9, ip, (size_t)(iend - ip), LL_base, LL_bits, LL_defnorm, 6, c->fseValid);
        if (h < 0) return -E_CORRUPTION;
        ip += h;
        h = build_seq_table(&c->of, (modes >> 4) & 3, 31, 8, ip, (size_t)(iend - ip), OF_base, OF_bits, OF_defnorm, 5, c->fseValid);
        if (h < 0) return -E_CORRUPTION;
        ip += h;
        h = build_seq_table(&c->ml, (modes >> 2) & 3, 52, 9, ip, (size_t)(iend - ip), ML_base, ML_bits, ML_defnorm, 6, c->fseValid);
        if (h < 0) return -E_CORRUPTION;
        ip += h;
    }
    /* sequences */
    u8* op = dst;
    u8* oend = dst + cap;
    const u8* litEnd = litPtr + litSize;
    if (nbSeq) {
        bstream b;
        c->fseValid = 1;
        if (bs_init(&b, ip, (size_t)(iend - ip)) < 0) return -E_CORRUPTION;
        u32 sLL = (u32)bs_read(&b, c->ll.log), sOF = (u32)bs_read(&b, c->of.log), sML = (u32)bs_read(&b, c->ml.log);
        u64 rep0 = c->rep[0], rep1 = c->rep[1], rep2 = c->rep[2];
        i64 err = 0;
        for (int n = 0; n < nbSeq; n++) {
            fse_entry eLL = c->ll.e[sLL], eML = c->ml.e[sML], eOF = c->of.e[sOF];
            u64 offset;
            if (eOF.nbAddBits > 1) {
                offset = eOF.baseValue + bs_read(&b, eOF.nbAddBits);
                rep2 = rep1; rep1 = rep0; rep0 = offset;
            } else {
                u32 ll0 = (eLL.baseValue == 0);
                if (eOF.nbAddBits == 0) {
                    if (!ll0) offset = rep0;
                    else { offset = rep1; rep1 = rep0; rep0 = offset; }
                } else {
                    u32 idx = eOF.baseValue + ll0 + (u32)bs_read(&b, 1);
                    u64 t = (idx == 3) ? rep0 - 1 : (idx == 1 ? rep1 : rep2);
                    t += !t;
                    if (idx != 1) rep2 = rep1;
                    rep1 = rep0; rep0 = offset = t;
                }
            }
            u64 ml = eML.baseValue + (eML.nbAddBits ? bs_read(&b, eML.nbAddBits) : 0);
            u64 ll = eLL.baseValue + (eLL.nbAddBits ? bs_read(&b, eLL.nbAddBits) : 0);
            /* state updates happen after every sequence, the last one included (reads may run dry) */
            sLL = eLL.nextState + (u32)bs_read(&b, eLL.nbBits);
            sML = eML.nextState + (u32)bs_read(&b, eML.nbBits);
            sOF = eOF.nextState + (u32)bs_read(&b, eOF.nbBits);
            if (n == nbSeq - 1) {
                /* undo the last update's consumption: the reference checks the stream right after
                 * the final sequence's value bits; the trailing state update is harmless there
                 * because an over-read is not an error in zstd_decompress_block.c:1102 */
                b.bits += eLL.nbBits + eML.nbBits + eOF.nbBits;
            }
            if (err) continue;
            /* execute */
            if (ll > (u64)(litEnd - litPtr)) { err = -E_CORRUPTION; continue; }
            if (ll + ml > (u64)(oend - op)) { err = -E_DST_TOO_SMALL; continue; }
            memcpy(op, litPtr, ll); op += ll; litPtr += ll;
            if (offset > (u64)(op - frameStart)) { err = -E_CORRUPTION; continue; }
            { const u8* m = op - offset; for (u64 i = 0; i < ml; i++) op[i] = m[i]; }
            op += ml;
        }
        if (err) return err;
        if (b.bits != 0) return -E_CORRUPTION; /* under- or over-consumed (over-read is flagged here) */
        c->rep[0] = (u32)rep0; c->rep[1] = (u32)rep1; c->rep[2] = (u32)rep2;
    }
    {
        size_t last = (size_t)(litEnd - litPtr);
        if (last > (size_t)(oend - op)) return -E_DST_TOO_SMALL;
        memcpy(op, litPtr, last);
        op += last;
    }
    return op - dst;
}

/* One zstd frame at *srcp. Advances *srcp/*lenp past it. */
static i64 decode_frame(frame_ctx* c, u8* dst, size_t cap, const u8** srcp, size_t* lenp) {
    const u8* ip = *srcp;
    size_t rem = *lenp;
    if (rem < 6 + 3) return -E_SRC_WRONG;
    if (rd32(ip) != 0xFD2FB528u) return -E_PREFIX_UNKNOWN;
    u8 fhd = ip[4];
    int didCode = fhd & 3, checksum = (fhd >> 2) & 1, single = (fhd >> 5) & 1, fcsId = fhd >> 6;
    static const int didSz[4] = {0, 1, 2, 4}, fcsSz[4] = {0, 2, 4, 8};
    size_t hsz = 5 + !single + (size_t)didSz[didCode] + (size_t)fcsSz[fcsId] + (size_t)(single && !fcsId);
    if (rem < hsz + 3) return -E_SRC_WRONG;
    if (fhd & 8) return -E_FRAMEPARAM_UNSUPPORTED;
    size_t pos = 5;
    u64 fcs = ~0ULL;
    if (!single) {
        u32 wl = (ip[pos++] >> 3) + 10;
        if (wl > 31) return -E_WINDOW_TOO_LARGE;
    }
    u32 dictID = 0;
    if (didCode == 1) dictID = ip[pos]; else if (didCode == 2) dictID = rd16(ip + pos); else if (didCode == 3) dictID = rd32(ip + pos);
    pos += (size_t)didSz[didCode];
    if (fcsId == 0) { if (single) fcs = ip[pos]; }
    else if (fcsId == 1) fcs = rd16(ip + pos) + 256;
    else if (fcsId == 2) fcs = rd32(ip + pos);
    else fcs = rd64(ip + pos);
    if (dictID) return -E_DICT_WRONG;
    ip += hsz; rem -= hsz;
    /* per-frame reset (zstd_decompress.c:1158-1179) */
    c->rep[0] = 1; c->rep[1] = 4; c->rep[2] = 8;
    c->fseValid = 0; c->huf.valid = 0;
    u8* op = dst;
    u8* oend = dst + cap;
    for (;;) {
        if (rem < 3) return -E_SRC_WRONG;
        u32 bh = rd24(ip);
        int last = bh & 1, type = (bh >> 1) & 3;
        size_t bsz = bh >> 3, csz = (type == 1) ? 1 : bsz;
        if (type == 3) return -E_CORRUPTION;
        ip += 3; rem -= 3;
        if (csz > rem) return -E_SRC_WRONG;
        i64 d;
        if (type == 2) d = decode_block(c, op, (size_t)(oend - op), dst, ip, csz);
        else if (type == 0) { if (bsz > (size_t)(oend - op)) return -E_DST_TOO_SMALL; memcpy(op, ip, bsz); d = (i64)bsz; }
        else { if (bsz > (size_t)(oend - op)) return -E_DST_TOO_SMALL; memset(op, ip[0], bsz); d = (i64)bsz; }
        if (d < 0) return d;
        op += d; ip += csz; rem -= csz;
        if (last) break;
    }
    if (fcs != ~0ULL && (u64)(op - dst) != fcs) return -E_CORRUPTION;
    if (checksum) {
        if (rem < 4) return -E_CHECKSUM_WRONG;
        if (rd32(ip) != (u32)zra_oracle_xxh64(dst, (size_t)(op - dst), 0)) return -E_CHECKSUM_WRONG;
        ip += 4; rem -= 4;
    }
    *srcp = ip; *lenp = rem;
    return op - dst;
}

/* Concatenated frames incl. skippable ones == ZSTD_decompressDCtx (zstd_decompress.c:694-785). */
ORACLE_API i64 zra_oracle_zstd_decompress(void* dstv, size_t cap, const void* srcv, size_t len) {
    u8* dst = (u8*)dstv;
    const u8* src = (const u8*)srcv;
    frame_ctx* c = (frame_ctx*)malloc(sizeof(frame_ctx));
    if (!c) return -64;
    c->lit = (u8*)malloc(BLOCKSIZE_MAX + 32);
    u8* op = dst;
    int more = 0;
    i64 rc = 0;
    while (len >= 5) {
        u32 magic = rd32(src);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (len < 8) { rc = -E_SRC_WRONG; break; }
            u64 skip = (u64)rd32(src + 4) + 8;
            if (skip > len) { rc = -E_SRC_WRONG; break; }
            src += skip; len -= skip;
            continue;
        }
        i64 d = decode_frame(c, op, cap - (size_t)(op - dst), &src, &len);
        if (d == -E_PREFIX_UNKNOWN && more) d = -E_SRC_WRONG;
        if (d < 0) { rc = d; break; }
        op += d;
        more = 1;
    }
    if (!rc && len) rc = -E_SRC_WRONG;
    free(c->lit); free(c);
    return rc ? rc : (i64)(op - dst);
}

/* --------------------------------------------------------------- ZRA layer */
enum { ZRA_OK = 0, ZRA_ZSTD = 1, ZRA_VERSION_LOW = 2, ZRA_HEADER_INVALID = 3, ZRA_HEADER_INCOMPLETE = 4,
       ZRA_OOB = 5, ZRA_OUT_TOO_SMALL = 6, ZRA_TOO_LARGE = 7, ZRA_FRAME_MISMATCH = 8 };

typedef struct {
    u32 version, size, frameSize, metaOffset, metaSize, seekTableOffset, seekTableSize, tableSize, hash;
    u64 uncompressedSize;
} zra_oracle_header;

/* ZSTD_COMPRESSBOUND (zstd/lib/zstd.h:174) */
ORACLE_API u64 zra_oracle_compress_bound(u64 s) {
    return s + (s >> 8) + ((s < (128u << 10)) ? (((128u << 10) - s) >> 11) : 0);
}

/* zra::GetOutputBufferSize (source/zra.cpp:189-192) */
ORACLE_API u64 zra_oracle_output_buffer_size(u64 inputSize, u32 frameSize, u32 metaSize) {
    u32 table = (u32)(inputSize / frameSize) + ((inputSize % frameSize) ? 2 : 1);
    return 38 + (u64)metaSize + 5ull * table + zra_oracle_compress_bound(frameSize) * (table - 1);
}

/* Header::Header(BufferView) (source/zra.cpp:141-171). Returns a ZRA status code. */
ORACLE_API int zra_oracle_parse_header(const void* buf, size_t n, zra_oracle_header* h) {
    const u8* p = (const u8*)buf;
    if (38 >= n) return ZRA_OOB; /* the reference's read lambda rejects offset+size >= buffer.size */
    if (rd32(p + 8) != 0x3041525Au || rd16(p + 12) > 1) return ZRA_HEADER_INVALID;
    h->version = rd16(p + 12);
    h->size = rd32(p + 4) + 8;
    h->hash = rd32(p + 14);
    h->uncompressedSize = rd64(p + 18);
    h->tableSize = rd32(p + 26);
    h->frameSize = rd32(p + 30);
    h->metaOffset = 38;
    h->metaSize = rd32(p + 34);
    h->seekTableOffset = 38 + h->metaSize;
    h->seekTableSize = h->tableSize * 5;
    if (h->version != 1) return ZRA_VERSION_LOW;
    if (n < h->size) return ZRA_OOB;
    return ZRA_OK;
}

static u64 entry40(const u8* p) { return (u64)rd32(p) | ((u64)p[4] << 32); }

/* Serialises FixedHeader ‖ meta ‖ Entry[frames+1] with the CRC filled in (source/zra.cpp:111-134).
 * `offsets` has frames+1 values (last = total compressed size). Returns bytes written. */
ORACLE_API u64 zra_oracle_build_header(void* out, u64 uncompressedSize, u32 frameSize, const void* meta, u32 metaSize,
                                       const u64* offsets, u32 tableSize) {
    u8* p = (u8*)out;
    u32 headerSize = 38 + metaSize + 5 * tableSize - 8;
    u32 f32[3] = {0x184D2A50u, headerSize, 0x3041525Au};
    for (int i = 0; i < 3; i++) for (int k = 0; k < 4; k++) p[4 * i + k] = (u8)(f32[i] >> (8 * k));
    p[12] = 1; p[13] = 0;
    memset(p + 14, 0, 4);
    for (int k = 0; k < 8; k++) p[18 + k] = (u8)(uncompressedSize >> (8 * k));
    u32 g32[3] = {tableSize, frameSize, metaSize};
    for (int i = 0; i < 3; i++) for (int k = 0; k < 4; k++) p[26 + 4 * i + k] = (u8)(g32[i] >> (8 * k));
    if (metaSize) memcpy(p + 38, meta, metaSize);
    u8* t = p + 38 + metaSize;
    for (u32 i = 0; i < tableSize; i++) for (int k = 0; k < 5; k++) t[5 * i + k] = (u8)(offsets[i] >> (8 * k));
    u32 crc = zra_oracle_crc32(p, 14, 0);
    crc = zra_oracle_crc32(p + 18, 20, crc);
    crc = zra_oracle_crc32(p + 38, metaSize + 5 * tableSize, crc);
    for (int k = 0; k < 4; k++) p[14 + k] = (u8)(crc >> (8 * k));
    return 38ull + metaSize + 5ull * tableSize;
}

/* Recomputes the header CRC of an archive (FixedHeader::CalculateHash, source/zra.cpp:128-133). */
ORACLE_API u32 zra_oracle_header_crc(const void* buf, size_t n) {
    const u8* p = (const u8*)buf;
    if (n < 38) return 0;
    u32 total = rd32(p + 4) + 8;
    if (total > n || total < 38) return 0;
    u32 crc = zra_oracle_crc32(p, 14, 0);
    crc = zra_oracle_crc32(p + 18, 20, crc);
    return zra_oracle_crc32(p + 38, total - 38, crc);
}

/* zra::DecompressBuffer (source/zra.cpp:243-250). status[0]=zra code, status[1]=zstd code. */
ORACLE_API i64 zra_oracle_decompress_buffer(const void* in, size_t n, void* out, size_t outCap, int* status) {
    zra_oracle_header h;
    status[1] = 0;
    if ((status[0] = zra_oracle_parse_header(in, n, &h)) != ZRA_OK) return -1;
    if (outCap < h.uncompressedSize) { status[0] = ZRA_OUT_TOO_SMALL; return -1; }
    i64 r = zra_oracle_zstd_decompress(out, outCap, (const u8*)in + h.size, n - h.size);
    if (r < 0) { status[0] = ZRA_ZSTD; status[1] = (int)-r; return -1; }
    return r;
}

/* zra::DecompressRA (source/zra.cpp:258-296) when inMemoryQuirk != 0 (bound check `>=`),
 * zra::Decompressor::Decompress (source/zra.cpp:369-413) otherwise (bound check `>`). */
ORACLE_API i64 zra_oracle_decompress_ra(const void* in, size_t n, void* outv, size_t outCap, u64 offset, u64 size,
                                        int inMemoryQuirk, int* status) {
    zra_oracle_header h;
    const u8* base = (const u8*)in;
    u8* out = (u8*)outv;
    status[1] = 0;
    if ((status[0] = zra_oracle_parse_header(in, n, &h)) != ZRA_OK) return -1;
    if (inMemoryQuirk ? (offset + size >= h.uncompressedSize) : (offset + size > h.uncompressedSize)) { status[0] = ZRA_OOB; return -1; }
    if (outCap < size) { status[0] = ZRA_OUT_TOO_SMALL; return -1; }
    u64 q = offset / h.frameSize, r = offset % h.frameSize;
    u64 q2 = (r + size) / h.frameSize, r2 = (r + size) % h.frameSize;
    const u8* table = base + h.seekTableOffset;
    u64 first = q, last = q + q2 + (r2 ? 1 : 0);
    const u8* contents = base + h.size;
    u8* tmp = (u8*)malloc(h.frameSize ? h.frameSize : 1);
    u64 done = 0;
    i64 d = 0;
    if (r) {
        u64 a = entry40(table + 5 * first), b = entry40(table + 5 * (first + 1));
        d = zra_oracle_zstd_decompress(tmp, h.frameSize, contents + a, b - a);
        if (d < 0) goto fail;
        u64 m = h.frameSize - r; if (size < m) m = size;
        memcpy(out, tmp + r, m);
        done += m; first++;
    }
    if (done < size) {
        u64 a = entry40(table + 5 * first), b = entry40(table + 5 * (r2 ? last - 1 : last));
        d = zra_oracle_zstd_decompress(out + done, outCap - done, contents + a, b - a);
        if (d < 0) goto fail;
        done += (u64)d;
    }
    if (done < size && r2) {
        u64 a = entry40(table + 5 * (last - 1)), b = entry40(table + 5 * last);
        d = zra_oracle_zstd_decompress(tmp, h.frameSize, contents + a, b - a);
        if (d < 0) goto fail;
        memcpy(out + done, tmp, size - done);
    }
    free(tmp);
    return (i64)size;
fail:
    free(tmp);
    status[0] = ZRA_ZSTD; status[1] = (int)-d;
    return -1;
}
