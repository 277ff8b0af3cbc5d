/* oracle/zstd_oracle.c - TEST INFRASTRUCTURE ONLY.  See vbz_oracle.h.
 *
 * Stage 2 of the reference is the third-party libzstd (call sites vbz/vbz.cpp:109,194-204,
 * 236-240,263-272; pinned zstd/1.4.8 by conan, CMakeLists.txt:93; source NOT under
 * /root/reference).  This file restates the published Zstandard format (RFC 8878) as a plain,
 * single-threaded, table-driven decoder; it shares no code with the product's CUDA decoder.
 * It is pinned against (a) the ten frames stored in the reference's fast5 fixtures and
 * (b) frames produced by the image's libzstd 1.5.5 at many levels (tests/test_oracle.py).
 *
 * The encode direction is not byte-pinned by the reference: vbzo_libzstd_*
 * call the image's runtime libzstd.so.1 through dlopen.
 */
#include "vbz_oracle.h"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#define ZERR ((size_t)-1)
#define BLOCK_MAX (128u * 1024u)

/* ---- libzstd through dlopen ----------------------------------------------- */
typedef size_t (*zc_fn)(void*, size_t, const void*, size_t, int);
typedef size_t (*zd_fn)(void*, size_t, const void*, size_t);
typedef unsigned (*ze_fn)(size_t);
static void* g_lib;
static zc_fn g_zc;
static zd_fn g_zd;
static ze_fn g_ze;
static int load_libzstd(void)
{
    if (g_lib) return 1;
    const char* names[] = {"libzstd.so.1", "/usr/lib/x86_64-linux-gnu/libzstd.so.1", "libzstd.so"};
    for (int i = 0; i < 3 && !g_lib; ++i) g_lib = dlopen(names[i], RTLD_NOW);
    if (!g_lib) return 0;
    g_zc = (zc_fn)dlsym(g_lib, "ZSTD_compress");
    g_zd = (zd_fn)dlsym(g_lib, "ZSTD_decompress");
    g_ze = (ze_fn)dlsym(g_lib, "ZSTD_isError");
    return g_zc && g_zd && g_ze;
}
size_t vbzo_libzstd_compress(void* dst, size_t cap, const void* src, size_t n, int level)
{
    if (!load_libzstd()) return ZERR;
    size_t r = g_zc(dst, cap, src, n, level);
    return g_ze(r) ? ZERR : r;
}
size_t vbzo_libzstd_decompress(void* dst, size_t cap, const void* src, size_t n)
{
    if (!load_libzstd()) return ZERR;
    size_t r = g_zd(dst, cap, src, n);
    return g_ze(r) ? ZERR : r;
}

/* ZSTD_COMPRESSBOUND (zstd.h): n + n/256 + (n < 128 KiB ? (128 KiB - n)/2048 : 0) */
size_t vbzo_zstd_compress_bound(size_t n)
{
    return n + (n >> 8) + (n < BLOCK_MAX ? ((BLOCK_MAX - n) >> 11) : 0);
}

/* ---- small helpers ---------------------------------------------------------- */
static int highbit(uint32_t v) { int r = 0; while (v >>= 1) ++r; return r; }
static uint32_t le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static uint32_t le24(const uint8_t* p) { return le16(p) | ((uint32_t)p[2] << 16); }
static uint32_t le32(const uint8_t* p) { return le24(p) | ((uint32_t)p[3] << 24); }
static uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

/* forward, LSB-first bit reader (FSE table descriptions) */
typedef struct { const uint8_t* p; size_t n; size_t bit; } FwdBits;
static uint32_t fwd_peek(const FwdBits* b, int nb)
{
    uint32_t v = 0;
    for (int i = 0; i < nb; ++i) {
        size_t pos = b->bit + (size_t)i;
        uint32_t bitv = (pos >> 3) < b->n ? (b->p[pos >> 3] >> (pos & 7)) & 1u : 0u;
        v |= bitv << i;
    }
    return v;
}

/* backward bit reader: bits are consumed from the top of a little-endian big integer whose
 * highest set bit is a padding marker.  Reading below bit 0 yields zeros; `pos` going negative
 * records the overrun. */
typedef struct { const uint8_t* p; long long pos; } BackBits;
static int back_init(BackBits* b, const uint8_t* p, size_t n)
{
    if (n == 0 || p[n - 1] == 0) return 0;
    b->p = p;
    b->pos = (long long)(n - 1) * 8 + highbit(p[n - 1]);
    return 1;
}
static uint32_t back_peek(const BackBits* b, int nb) /* next nb bits, first-read bit = MSB */
{
    uint32_t v = 0;
    for (int i = 0; i < nb; ++i) {
        long long pos = b->pos - 1 - i;
        uint32_t bitv = pos >= 0 ? (b->p[pos >> 3] >> (pos & 7)) & 1u : 0u;
        v = (v << 1) | bitv;
    }
    return v;
}
static uint32_t back_read(BackBits* b, int nb) { uint32_t v = back_peek(b, nb); b->pos -= nb; return v; }

/* ---- FSE ----------------------------------------------------------------------- */
typedef struct { uint8_t sym; uint8_t nb; uint16_t base; } FseCell;
typedef struct { int al; FseCell cell[512]; } FseTable;

/* table description -> normalised counts.  Returns bytes consumed or 0 on error. */
static size_t fse_read_ncount(const uint8_t* p, size_t n, int max_al, int max_sym, int16_t* norm, int* nsym, int* al_out)
{
    FwdBits b = {p, n, 0};
    if (n == 0) return 0;
    int al = (int)fwd_peek(&b, 4) + 5;
    b.bit += 4;
    if (al > max_al) return 0;
    int remaining = (1 << al) + 1, thr = 1 << al, nb = al + 1, sym = 0, prev0 = 0;
    while (remaining > 1 && sym <= max_sym) {
        if (prev0) {
            for (;;) {
                uint32_t r = fwd_peek(&b, 2);
                b.bit += 2;
                for (uint32_t k = 0; k < r; ++k) { if (sym > max_sym) return 0; norm[sym++] = 0; }
                if (r != 3) break;
            }
            if (sym > max_sym) return 0;
        }
        int max = (2 * thr - 1) - remaining, c;
        uint32_t low = fwd_peek(&b, nb - 1);
        if ((int)low < max) { c = (int)low; b.bit += (size_t)(nb - 1); }
        else { c = (int)fwd_peek(&b, nb); if (c >= thr) c -= max; b.bit += (size_t)nb; }
        c -= 1;
        remaining -= c < 0 ? -c : c;
        norm[sym++] = (int16_t)c;
        prev0 = (c == 0);
        while (remaining < thr) { nb--; thr >>= 1; }
    }
    if (remaining != 1) return 0;
    size_t used = (b.bit + 7) / 8;
    if (used > n) return 0;
    *nsym = sym;
    *al_out = al;
    return used;
}

static int fse_build(FseTable* t, const int16_t* norm, int nsym, int al)
{
    int T = 1 << al, high = T - 1;
    uint16_t next[256];
    t->al = al;
    for (int s = 0; s < nsym; ++s) {
        if (norm[s] == -1) { t->cell[high--].sym = (uint8_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
    }
    int step = (T >> 1) + (T >> 3) + 3, pos = 0;
    for (int s = 0; s < nsym; ++s)
        for (int i = 0; i < norm[s]; ++i) {
            t->cell[pos].sym = (uint8_t)s;
            do { pos = (pos + step) & (T - 1); } while (pos > high);
        }
    if (pos != 0) return 0;
    for (int u = 0; u < T; ++u) {
        uint32_t x = next[t->cell[u].sym]++;
        int nb = al - highbit(x);
        t->cell[u].nb = (uint8_t)nb;
        t->cell[u].base = (uint16_t)((x << nb) - (uint32_t)T);
    }
    return 1;
}
static void fse_rle(FseTable* t, int sym) { t->al = 0; t->cell[0].sym = (uint8_t)sym; t->cell[0].nb = 0; t->cell[0].base = 0; }

/* ---- Huffman ---------------------------------------------------------------------- */
typedef struct { int max_bits; uint8_t sym[4096]; uint8_t nb[4096]; } HufTable;

/* tree description -> decode table.  Returns bytes consumed or 0 on error. */
static size_t huf_read_table(HufTable* h, const uint8_t* p, size_t n)
{
    uint8_t w[256];
    int nw = 0;
    size_t used;
    if (n < 1) return 0;
    int hb = p[0];
    if (hb >= 128) {
        nw = hb - 127;
        used = 1 + (size_t)(nw + 1) / 2;
        if (used > n) return 0;
        for (int i = 0; i < nw; ++i) w[i] = (i & 1) ? (p[1 + i / 2] & 0xF) : (p[1 + i / 2] >> 4);
    } else {
        used = 1 + (size_t)hb;
        if (used > n || hb == 0) return 0;
        int16_t norm[256];
        int nsym, al;
        size_t hdr = fse_read_ncount(p + 1, (size_t)hb, 6, 255, norm, &nsym, &al);
        if (!hdr) return 0;
        FseTable t;
        if (!fse_build(&t, norm, nsym, al)) return 0;
        BackBits b;
        if (!back_init(&b, p + 1 + hdr, (size_t)hb - hdr)) return 0;
        uint32_t s1 = back_read(&b, al), s2 = back_read(&b, al);
        for (;;) { /* two interleaved states; stop one symbol after the stream overruns */
            if (nw > 253) return 0;
            w[nw++] = t.cell[s1].sym;
            s1 = t.cell[s1].base + back_read(&b, t.cell[s1].nb);
            if (b.pos < 0) { w[nw++] = t.cell[s2].sym; break; }
            if (nw > 253) return 0;
            w[nw++] = t.cell[s2].sym;
            s2 = t.cell[s2].base + back_read(&b, t.cell[s2].nb);
            if (b.pos < 0) { w[nw++] = t.cell[s1].sym; break; }
        }
    }
    uint32_t total = 0, rank[16] = {0};
    for (int i = 0; i < nw; ++i) {
        if (w[i] > 12) return 0;
        rank[w[i]]++;
        total += (1u << w[i]) >> 1;
    }
    if (total == 0) return 0;
    int max_bits = highbit(total) + 1;
    if (max_bits > 12) return 0;
    uint32_t rest = (1u << max_bits) - total;
    if ((rest & (rest - 1)) != 0) return 0;
    w[nw] = (uint8_t)(highbit(rest) + 1);
    rank[w[nw]]++;
    nw++;
    if (rank[1] < 2 || (rank[1] & 1)) return 0;
    /* cells: weight 1 (longest codes) first, symbols ascending within a weight */
    uint32_t start[16], acc = 0;
    for (int r = 1; r <= max_bits; ++r) { start[r] = acc; acc += rank[r] << (r - 1); }
    for (int s = 0; s < nw; ++s) {
        int r = w[s];
        if (!r) continue;
        uint32_t len = (1u << r) >> 1;
        for (uint32_t k = 0; k < len; ++k) { h->sym[start[r] + k] = (uint8_t)s; h->nb[start[r] + k] = (uint8_t)(max_bits + 1 - r); }
        start[r] += len;
    }
    h->max_bits = max_bits;
    return used;
}

static int huf_decode_stream(const HufTable* h, const uint8_t* p, size_t n, uint8_t* out, size_t count)
{
    BackBits b;
    if (!back_init(&b, p, n)) return 0;
    for (size_t i = 0; i < count; ++i) {
        uint32_t idx = back_peek(&b, h->max_bits);
        out[i] = h->sym[idx];
        b.pos -= h->nb[idx];
    }
    return b.pos == 0;
}

/* ---- frame state --------------------------------------------------------------------- */
typedef struct {
    HufTable huf; int huf_valid;
    FseTable ll, of, ml; int seq_valid;
    uint32_t rep[3];
    uint8_t lit[BLOCK_MAX + 32];
} Frame;

static const int16_t LL_DEF[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const int16_t ML_DEF[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static const int16_t OF_DEF[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
static const uint32_t LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
static const uint8_t LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const uint32_t ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
static const uint8_t ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

/* literals section -> f->lit[0..*lit_n).  Returns bytes consumed or 0 on error. */
static size_t decode_literals(Frame* f, const uint8_t* p, size_t n, size_t* lit_n)
{
    if (n < 1) return 0;
    int type = p[0] & 3, sf = (p[0] >> 2) & 3;
    size_t hdr, regen, comp = 0;
    int streams = 4;
    if (type < 2) {
        if (sf == 0 || sf == 2) { hdr = 1; regen = p[0] >> 3; }
        else if (sf == 1) { if (n < 2) return 0; hdr = 2; regen = le16(p) >> 4; }
        else { if (n < 3) return 0; hdr = 3; regen = le24(p) >> 4; }
        if (regen > BLOCK_MAX) return 0;
        if (type == 0) {
            if (hdr + regen > n) return 0;
            memcpy(f->lit, p + hdr, regen);
            *lit_n = regen;
            return hdr + regen;
        }
        if (hdr + 1 > n) return 0;
        memset(f->lit, p[hdr], regen);
        *lit_n = regen;
        return hdr + 1;
    }
    if (n < 3) return 0;
    if (sf <= 1) { hdr = 3; uint32_t v = le24(p); regen = (v >> 4) & 0x3FF; comp = (v >> 14) & 0x3FF; if (sf == 0) streams = 1; }
    else if (sf == 2) { if (n < 4) return 0; hdr = 4; uint32_t v = le32(p); regen = (v >> 4) & 0x3FFF; comp = (v >> 18) & 0x3FFF; }
    else { if (n < 5) return 0; hdr = 5; uint64_t v = (uint64_t)le32(p) | ((uint64_t)p[4] << 32); regen = (size_t)((v >> 4) & 0x3FFFF); comp = (size_t)((v >> 22) & 0x3FFFF); }
    if (regen > BLOCK_MAX || hdr + comp > n) return 0;
    if (streams == 4 && regen < 6) return 0; /* libzstd >= 1.5.4: MIN_LITERALS_FOR_4_STREAMS */
    const uint8_t* q = p + hdr;
    size_t left = comp;
    if (type == 2) {
        size_t t = huf_read_table(&f->huf, q, left);
        if (!t) return 0;
        f->huf_valid = 1;
        q += t;
        left -= t;
    } else if (!f->huf_valid) return 0;
    if (streams == 1) {
        if (!huf_decode_stream(&f->huf, q, left, f->lit, regen)) return 0;
    } else {
        if (left < 10) return 0;
        size_t s1 = le16(q), s2 = le16(q + 2), s3 = le16(q + 4);
        if (6 + s1 + s2 + s3 > left) return 0;
        size_t s4 = left - 6 - s1 - s2 - s3;
        size_t seg = (regen + 3) / 4;
        if (3 * seg > regen) return 0;
        const uint8_t* d = q + 6;
        if (!huf_decode_stream(&f->huf, d, s1, f->lit, seg)) return 0;
        if (!huf_decode_stream(&f->huf, d + s1, s2, f->lit + seg, seg)) return 0;
        if (!huf_decode_stream(&f->huf, d + s1 + s2, s3, f->lit + 2 * seg, seg)) return 0;
        if (!huf_decode_stream(&f->huf, d + s1 + s2 + s3, s4, f->lit + 3 * seg, regen - 3 * seg)) return 0;
    }
    *lit_n = regen;
    return hdr + comp;
}

/* one of the three sequence tables, by mode.  Returns bytes consumed (may be 0 for modes 0/3), ZERR on error. */
static size_t seq_table(FseTable* t, int mode, const uint8_t* p, size_t n, int max_al, int max_sym,
                        const int16_t* def, int def_n, int def_al, int have_prev)
{
    if (mode == 0) { if (!fse_build(t, def, def_n, def_al)) return ZERR; return 0; }
    if (mode == 1) { if (n < 1 || p[0] > max_sym) return ZERR; fse_rle(t, p[0]); return 1; }
    if (mode == 3) return have_prev ? 0 : ZERR;
    int16_t norm[64];
    int nsym, al;
    size_t used = fse_read_ncount(p, n, max_al, max_sym, norm, &nsym, &al);
    if (!used) return ZERR;
    if (!fse_build(t, norm, nsym, al)) return ZERR;
    return used;
}

/* one compressed block: literals + sequences -> out[*op..]; base = start of this frame's output */
static int decode_compressed_block(Frame* f, const uint8_t* p, size_t n, uint8_t* out, size_t cap, size_t* op)
{
    size_t lit_n = 0;
    size_t used = decode_literals(f, p, n, &lit_n);
    if (!used) return 0;
    p += used;
    n -= used;
    if (n < 1) return 0;
    uint32_t nseq = p[0];
    size_t h = 1;
    if (nseq == 0) {
        if (n != 1) return 0;
        if (*op + lit_n > cap) return 0;
        memcpy(out + *op, f->lit, lit_n);
        *op += lit_n;
        return 1;
    }
    if (nseq > 0x7F) {
        if (nseq == 0xFF) { if (n < 3) return 0; nseq = le16(p + 1) + 0x7F00; h = 3; }
        else { if (n < 2) return 0; nseq = ((nseq - 0x80) << 8) + p[1]; h = 2; }
    }
    if (h + 1 > n) return 0;
    int modes = p[h++];
    if (modes & 3) return 0; /* reserved bits (checked by libzstd 1.5.x) */
    size_t r;
    r = seq_table(&f->ll, modes >> 6, p + h, n - h, 9, 35, LL_DEF, 36, 6, f->seq_valid); if (r == ZERR) return 0; h += r;
    r = seq_table(&f->of, (modes >> 4) & 3, p + h, n - h, 8, 31, OF_DEF, 29, 5, f->seq_valid); if (r == ZERR) return 0; h += r;
    r = seq_table(&f->ml, (modes >> 2) & 3, p + h, n - h, 9, 52, ML_DEF, 53, 6, f->seq_valid); if (r == ZERR) return 0; h += r;
    f->seq_valid = 1;
    BackBits b;
    if (!back_init(&b, p + h, n - h)) return 0;
    uint32_t sl = back_read(&b, f->ll.al), so = back_read(&b, f->of.al), sm = back_read(&b, f->ml.al);
    if (b.pos < 0) return 0;
    size_t lp = 0;
    for (uint32_t i = 0; i < nseq; ++i) {
        int llc = f->ll.cell[sl].sym, ofc = f->of.cell[so].sym, mlc = f->ml.cell[sm].sym;
        if (llc > 35 || mlc > 52 || ofc > 31) return 0;
        uint32_t ofv = (ofc ? (1u << ofc) : 1u) + (ofc ? back_read(&b, ofc) : 0u);
        uint32_t mlen = ML_BASE[mlc] + back_read(&b, ML_BITS[mlc]);
        uint32_t llen = LL_BASE[llc] + back_read(&b, LL_BITS[llc]);
        uint32_t offset;
        if (ofv > 3) {
            offset = ofv - 3;
            f->rep[2] = f->rep[1]; f->rep[1] = f->rep[0]; f->rep[0] = offset;
        } else {
            uint32_t k = ofv + (llen == 0 ? 1u : 0u);
            if (k == 1) offset = f->rep[0];
            else {
                offset = (k == 4) ? f->rep[0] - 1 : f->rep[k - 1];
                if (offset == 0) offset = 1; /* libzstd: "0 is not valid; input is corrupted; force offset to 1" */
                if (k != 2) f->rep[2] = f->rep[1];
                f->rep[1] = f->rep[0];
                f->rep[0] = offset;
            }
        }
        if (i + 1 < nseq) {
            sl = f->ll.cell[sl].base + back_read(&b, f->ll.cell[sl].nb);
            sm = f->ml.cell[sm].base + back_read(&b, f->ml.cell[sm].nb);
            so = f->of.cell[so].base + back_read(&b, f->of.cell[so].nb);
        }
        if (b.pos < 0) return 0;
        if (lp + llen > lit_n) return 0;
        if (*op + llen + mlen > cap) return 0;
        memcpy(out + *op, f->lit + lp, llen);
        lp += llen;
        *op += llen;
        if (offset > *op) return 0;
        for (uint32_t k = 0; k < mlen; ++k) out[*op + k] = out[*op + k - offset];
        *op += mlen;
    }
    if (b.pos != 0) return 0;
    size_t rest = lit_n - lp;
    if (*op + rest > cap) return 0;
    memcpy(out + *op, f->lit + lp, rest);
    *op += rest;
    return 1;
}

/* XXH64 (seed 0): only the frame checksum needs it */
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t xxh64(const uint8_t* p, size_t n)
{
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P3 = 1609587929392839161ULL,
                   P4 = 9650029242287828579ULL, P5 = 2870177450012600261ULL;
    const uint8_t* e = p + n;
    uint64_t h;
    if (n >= 32) {
        uint64_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
        do {
            v1 = rotl64(v1 + le64(p) * P2, 31) * P1; p += 8;
            v2 = rotl64(v2 + le64(p) * P2, 31) * P1; p += 8;
            v3 = rotl64(v3 + le64(p) * P2, 31) * P1; p += 8;
            v4 = rotl64(v4 + le64(p) * P2, 31) * P1; p += 8;
        } while (p + 32 <= e);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = (h ^ (rotl64(v1 * P2, 31) * P1)) * P1 + P4;
        h = (h ^ (rotl64(v2 * P2, 31) * P1)) * P1 + P4;
        h = (h ^ (rotl64(v3 * P2, 31) * P1)) * P1 + P4;
        h = (h ^ (rotl64(v4 * P2, 31) * P1)) * P1 + P4;
    } else h = P5;
    h += (uint64_t)n;
    while (p + 8 <= e) { h ^= rotl64(le64(p) * P2, 31) * P1; h = rotl64(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= e) { h ^= (uint64_t)le32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
    while (p < e) { h ^= (*p++) * P5; h = rotl64(h, 11) * P1; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* frame header.  Returns header length, 0 on error / truncated input. */
static size_t frame_header(const uint8_t* p, size_t n, unsigned long long* fcs, int* has_fcs, int* checksum)
{
    if (n < 5 || le32(p) != 0xFD2FB528u) return 0;
    int fhd = p[4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
    if (fhd & 0x08) return 0; /* reserved bit */
    size_t h = 5;
    if (!single) {
        if (h + 1 > n) return 0;
        int wl = 10 + (p[h] >> 3);
        if (wl > 31) return 0;
        h++;
    }
    static const int did_bytes[4] = {0, 1, 2, 4};
    if (h + (size_t)did_bytes[did] > n) return 0;
    uint32_t dict_id = 0;
    for (int i = 0; i < did_bytes[did]; ++i) dict_id |= (uint32_t)p[h + (size_t)i] << (8 * i);
    h += (size_t)did_bytes[did];
    if (dict_id != 0) return 0; /* no dictionary is ever supplied on this path */
    int fb = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (h + (size_t)fb > n) return 0;
    *has_fcs = fb != 0;
    *fcs = 0;
    if (fb == 1) *fcs = p[h];
    else if (fb == 2) *fcs = le16(p + h) + 256;
    else if (fb == 4) *fcs = le32(p + h);
    else if (fb == 8) *fcs = le64(p + h);
    *checksum = (fhd >> 2) & 1;
    return h + (size_t)fb;
}

/* ZSTD_getFrameContentSize */
unsigned long long vbzo_zstd_frame_content_size(const void* src, size_t n)
{
    const uint8_t* p = (const uint8_t*)src;
    if (n >= 8 && (le32(p) & 0xFFFFFFF0u) == 0x184D2A50u) return 0; /* skippable frame */
    unsigned long long fcs;
    int has, ck;
    if (!frame_header(p, n, &fcs, &has, &ck)) return VBZO_ZSTD_CONTENTSIZE_ERROR;
    return has ? fcs : VBZO_ZSTD_CONTENTSIZE_UNKNOWN;
}

/* ZSTD_decompress: every frame in src, back to back */
size_t vbzo_zstd_decompress(void* dst, size_t cap, const void* src, size_t n)
{
    const uint8_t* p = (const uint8_t*)src;
    uint8_t* out = (uint8_t*)dst;
    size_t total = 0;
    Frame* f = (Frame*)malloc(sizeof(Frame));
    if (!f) return ZERR;
    while (n > 0) {
        if (n >= 8 && (le32(p) & 0xFFFFFFF0u) == 0x184D2A50u) {
            size_t sk = (size_t)le32(p + 4) + 8;
            if (sk > n) goto fail;
            p += sk; n -= sk;
            continue;
        }
        unsigned long long fcs;
        int has, ck;
        size_t h = frame_header(p, n, &fcs, &has, &ck);
        if (!h) goto fail;
        p += h; n -= h;
        f->huf_valid = 0; f->seq_valid = 0;
        f->rep[0] = 1; f->rep[1] = 4; f->rep[2] = 8;
        uint8_t* fout = out + total;
        size_t fcap = cap - total, op = 0;
        for (;;) {
            if (n < 3) goto fail;
            uint32_t bh = le24(p);
            int last = bh & 1, type = (bh >> 1) & 3;
            size_t bs = bh >> 3;
            p += 3; n -= 3;
            if (type == 0) {
                if (bs > n || op + bs > fcap) goto fail;
                memcpy(fout + op, p, bs);
                op += bs; p += bs; n -= bs;
            } else if (type == 1) {
                if (n < 1 || op + bs > fcap) goto fail;
                memset(fout + op, p[0], bs);
                op += bs; p += 1; n -= 1;
            } else if (type == 2) {
                if (bs > n || bs >= BLOCK_MAX) goto fail;
                if (!decode_compressed_block(f, p, bs, fout, fcap, &op)) goto fail;
                p += bs; n -= bs;
            } else goto fail;
            if (last) break;
        }
        if (has && op != fcs) goto fail;
        if (ck) {
            if (n < 4) goto fail;
            if ((uint32_t)xxh64(fout, op) != le32(p)) goto fail;
            p += 4; n -= 4;
        }
        total += op;
    }
    free(f);
    return total;
fail:
    free(f);
    return ZERR;
}
