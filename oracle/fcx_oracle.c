/*
 * fcx_oracle.c -- CPU restatement of my_compress.cpp's lz77 path (see fcx_oracle.h).
 * TEST INFRASTRUCTURE ONLY; never linked into the product library.
 * "mc:N" = /root/reference/my_compress.cpp line N (\n-based).
 */
#include "fcx_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* little helpers                                                      */
/* ------------------------------------------------------------------ */
static void put_u32(uint8_t *dst, uint32_t v) { memcpy(dst, &v, 4); } /* host LE, as mc:2138 */
static uint32_t get_u32(const uint8_t *src) { uint32_t v; memcpy(&v, src, 4); return v; }

static uint32_t match_cap(uint32_t n, uint32_t i)
{
    /* mc:1455,1459: look-ahead = min(258, n-i); usable match length = look-ahead - 1 */
    uint32_t rem = n - i;
    uint32_t look = rem < FCXO_LOOKAHEAD ? rem : FCXO_LOOKAHEAD;
    return look - 1;
}

static uint32_t lcp_cap(const uint8_t *a, const uint8_t *b, uint32_t cap)
{
    uint32_t k = 0;
    while (k < cap && a[k] == b[k]) k++;
    return k;
}

/* ------------------------------------------------------------------ */
/* A.1 match rule                                                      */
/* ------------------------------------------------------------------ */
void fcxo_longest_match(const uint8_t *x, uint32_t n, uint32_t i, uint32_t *p, uint32_t *l)
{
    *p = 0; *l = 0;
    if (i == 0 || i >= n) return;                 /* mc:1451-1453 */
    uint32_t cap = match_cap(n, i);
    if (cap < FCXO_MIN_MATCH) return;             /* mc:1460-1462 */
    uint32_t s0 = i > FCXO_WINDOW ? i - FCXO_WINDOW : 0; /* mc:1456 */
    uint32_t best = 0, best_s = 0;
    for (uint32_t s = s0; s < i; s++) {           /* ascending s + strict '>' => farthest wins ties */
        uint32_t m = lcp_cap(x + s, x + i, cap);
        if (m > best) { best = m; best_s = s; }
    }
    if (best >= FCXO_MIN_MATCH) { *p = i - best_s; *l = best; }
}

/* hash chains over exact 3-byte grams */
typedef struct { int32_t *head; int32_t *prev; } chains_t;
#define HBITS 16
static uint32_t gram_hash(const uint8_t *q)
{
    uint32_t g = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16);
    return (g * 2654435761u) >> (32 - HBITS);
}

static void chains_build(chains_t *ch, const uint8_t *x, uint32_t n)
{
    ch->head = (int32_t *)malloc(sizeof(int32_t) << HBITS);
    ch->prev = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    for (uint32_t h = 0; h < (1u << HBITS); h++) ch->head[h] = -1;
    for (uint32_t i = 0; i < n; i++) ch->prev[i] = -1;
    for (uint32_t i = 0; i + 3 <= n; i++) {
        uint32_t h = gram_hash(x + i);
        ch->prev[i] = ch->head[h];
        ch->head[h] = (int32_t)i;
    }
}
static void chains_free(chains_t *ch) { free(ch->head); free(ch->prev); }

static void chain_match(const chains_t *ch, const uint8_t *x, uint32_t n, uint32_t i, uint32_t *p, uint32_t *l)
{
    *p = 0; *l = 0;
    if (i == 0 || i >= n) return;
    uint32_t cap = match_cap(n, i);
    if (cap < FCXO_MIN_MATCH) return;
    uint32_t best = 0, best_s = 0;
    /* nearest-first walk, so '>=' lets a farther candidate of equal length win */
    for (int32_t s = ch->prev[i]; s >= 0 && i - (uint32_t)s <= FCXO_WINDOW; s = ch->prev[s]) {
        if (x[s] != x[i] || x[s + 1] != x[i + 1] || x[s + 2] != x[i + 2]) continue;
        if (best >= FCXO_MIN_MATCH && x[s + best - 1] != x[i + best - 1]) continue;
        uint32_t m = lcp_cap(x + s, x + i, cap);
        if (m >= FCXO_MIN_MATCH && m >= best) { best = m; best_s = (uint32_t)s; }
    }
    if (best >= FCXO_MIN_MATCH) { *p = i - best_s; *l = best; }
}

void fcxo_match_all(const uint8_t *x, uint32_t n, uint32_t *p, uint32_t *l)
{
    chains_t ch; chains_build(&ch, x, n);
    for (uint32_t i = 0; i < n; i++) chain_match(&ch, x, n, i, &p[i], &l[i]);
    chains_free(&ch);
}

/* ------------------------------------------------------------------ */
/* A.2 greedy parse (mc:1683-1703)                                     */
/* ------------------------------------------------------------------ */
uint32_t fcxo_lz77_parse(const uint8_t *x, uint32_t n, uint32_t *p, uint32_t *l, uint8_t *c)
{
    chains_t ch; chains_build(&ch, x, n);
    uint32_t T = 0, cur = 0;
    while (cur < n) {
        uint32_t pp, ll;
        chain_match(&ch, x, n, cur, &pp, &ll);
        p[T] = pp; l[T] = ll; c[T] = x[cur + ll];   /* mc:1489: c = byte after the match */
        T++;
        cur += ll + 1;                              /* mc:1702 */
    }
    chains_free(&ch);
    return T;
}

/* ------------------------------------------------------------------ */
/* A.3 Huffman tree (mc:535-617)                                       */
/* ------------------------------------------------------------------ */
uint32_t fcxo_huffman_tree(const uint32_t w[256], uint32_t child[510])
{
    /* leaves sorted by weight, stable in symbol order (merge sort with '<=', mc:465) */
    uint32_t lw[256], ls[256], R = 0;
    for (uint32_t s = 0; s < 256; s++)
        if (w[s] > 0) {
            uint32_t k = R++;
            while (k > 0 && lw[k - 1] > w[s]) { lw[k] = lw[k - 1]; ls[k] = ls[k - 1]; k--; } /* stable insertion */
            lw[k] = w[s]; ls[k] = s;
        }
    memset(child, 0, 510 * sizeof(uint32_t));
    if (R < 2) return R;
    /* two-queue merge.  The reference re-inserts a merged node AFTER every entry
     * whose weight is <= it (strict '<' at mc:588), i.e. on equal weight a leaf or an
     * older internal node is taken first. */
    uint64_t iw[255];
    uint32_t base = 512 - R;                         /* mc:569: first internal node index */
    uint32_t qa = 0, qb = 0, made = 0;
    while ((R - qa) + (made - qb) > 1) {
        uint32_t pick[2];
        uint64_t sum = 0;
        for (int k = 0; k < 2; k++) {
            int take_leaf;
            if (qa >= R) take_leaf = 0;
            else if (qb >= made) take_leaf = 1;
            else take_leaf = (uint64_t)lw[qa] <= iw[qb];
            if (take_leaf) { pick[k] = ls[qa]; sum += lw[qa]; qa++; }
            else { pick[k] = base + qb; sum += iw[qb]; qb++; }
        }
        child[2 * made] = pick[0];                   /* first popped = left = bit 0 */
        child[2 * made + 1] = pick[1];
        iw[made] = sum;
        made++;
    }
    return R;
}

uint32_t fcxo_huffman_codes(uint32_t R, const uint32_t child[510], uint64_t bits[256], uint8_t len[256])
{
    memset(bits, 0, 256 * sizeof(uint64_t));
    memset(len, 0, 256);
    if (R < 2) return 0;                             /* mc:859: one leaf => no bits at all */
    uint32_t base = 512 - R, nint = R - 1, maxlen = 0;
    uint64_t ib[255]; uint32_t il[255];
    ib[nint - 1] = 0; il[nint - 1] = 0;              /* root is the last internal node */
    for (int32_t j = (int32_t)nint - 1; j >= 0; j--) {
        for (uint32_t side = 0; side < 2; side++) {
            uint32_t ci = child[2 * j + side];
            uint64_t b = ib[j] | ((uint64_t)side << il[j]);
            uint32_t d = il[j] + 1;
            if (ci < 256) { bits[ci] = b; len[ci] = (uint8_t)d; if (d > maxlen) maxlen = d; }
            else { ib[ci - base] = b; il[ci - base] = d; }
        }
    }
    return maxlen;
}

/* ------------------------------------------------------------------ */
/* A.4 Huffman section (mc:987-1104)                                   */
/* ------------------------------------------------------------------ */
/* x9: the opt-in FCX9 format (include/fcx.h) -- a section without tree nodes names its one symbol after the zero
 * count.  This is the only place the two payload formats differ on the encoder side. */
static uint32_t huffman_section_fmt(const uint8_t *src, uint32_t S, uint8_t *out, int x9)
{
    if (!src || !out || S == 0) return 0;            /* mc:989-990 */
    uint32_t w[256] = {0};
    for (uint32_t i = 0; i < S; i++) w[src[i]]++;
    uint32_t child[510];
    uint32_t R = fcxo_huffman_tree(w, child);
    uint32_t Tn = R > 1 ? R - 1 : 0;                 /* mc:1013-1019 */
    uint32_t mapB = (2 * Tn + 7) / 8;                /* mc:1024-1026 */
    uint8_t *q = out;
    *q++ = (uint8_t)Tn;
    if (x9 && Tn == 0) *q++ = src[0];                /* FCX9: the symbol FCX7 drops */
    memset(q, 0, mapB);
    for (uint32_t j = 0; j < Tn; j++)
        for (uint32_t side = 0; side < 2; side++)
            if (child[2 * j + side] >= 256) q[(2 * j + side) >> 3] |= (uint8_t)(1u << ((2 * j + side) & 7));
    q += mapB;
    for (uint32_t j = 0; j < Tn; j++)
        for (uint32_t side = 0; side < 2; side++) {
            uint32_t ci = child[2 * j + side];
            *q++ = (uint8_t)(ci >= 256 ? ci - 256 : ci);  /* mc:1035,1045 */
        }
    uint64_t bits[256]; uint8_t len[256];
    fcxo_huffman_codes(R, child, bits, len);
    uint64_t total = 0;
    for (uint32_t s = 0; s < 256; s++) total += (uint64_t)w[s] * len[s];
    uint32_t W = (uint32_t)((total + 31) / 32);      /* mc:895-923: flush on 32 bits + tail */
    put_u32(q, W); q += 4;
    memset(q, 0, (size_t)W * 4);
    uint64_t pos = 0;
    for (uint32_t i = 0; i < S; i++) {
        uint64_t b = bits[src[i]]; uint32_t n = len[src[i]];
        for (uint32_t k = 0; k < n; k++, pos++)
            if ((b >> k) & 1) q[pos >> 3] |= (uint8_t)(1u << (pos & 7));
    }
    q += (size_t)W * 4;
    return (uint32_t)(q - out);
}

uint32_t fcxo_huffman_section(const uint8_t *src, uint32_t S, uint8_t *out) { return huffman_section_fmt(src, S, out, 0); }

/* ------------------------------------------------------------------ */
/* A.5 Golomb-Rice (mc:258-304), p packing (mc:1292-1313)              */
/* ------------------------------------------------------------------ */
uint32_t fcxo_golomb_encode(const uint32_t *v, uint32_t cnt, uint32_t *words)
{
    uint64_t pos = 0;
    uint64_t total = 0;
    for (uint32_t i = 0; i < cnt; i++) total += (v[i] >> 2) + 3;
    uint32_t G = (uint32_t)((total + 31) / 32);
    memset(words, 0, (size_t)G * 4);
    for (uint32_t i = 0; i < cnt; i++) {
        uint32_t q = v[i] >> 2, r = v[i] & 3;
        for (uint32_t k = 0; k < q; k++, pos++) words[pos >> 5] |= 1u << (pos & 31);
        pos++;                                        /* the 0 separator */
        if (r & 1) words[pos >> 5] |= 1u << (pos & 31);
        pos++;
        if (r & 2) words[pos >> 5] |= 1u << (pos & 31);
        pos++;
    }
    return G;
}

uint32_t fcxo_combine_bits(const uint32_t *v, uint32_t cnt, uint32_t bits, uint8_t *out)
{
    uint32_t size = (bits * cnt) / 8 + 1;
    memset(out, 0, size);
    uint64_t pos = 0;
    for (uint32_t i = 0; i < cnt; i++)
        for (uint32_t k = 0; k < bits; k++, pos++)
            if ((v[i] >> k) & 1) out[pos >> 3] |= (uint8_t)(1u << (pos & 7));
    return size;
}

/* ------------------------------------------------------------------ */
/* A.6 block payload (mc:2115-2253) and container (mc:4073-4136)       */
/* ------------------------------------------------------------------ */
static uint32_t serialize_block_fmt(const uint32_t *p, const uint32_t *l, const uint8_t *c, uint32_t T, uint8_t *out, int x9)
{
    uint8_t *q = out;
    put_u32(q, T); q += 4;                                         /* mc:2137-2139 */
    /* literal-flag bitmap (mc:2073-2113) */
    uint32_t mapB = (T + 7) / 8;
    uint8_t *bm = (uint8_t *)calloc(mapB ? mapB : 1, 1);
    for (uint32_t t = 0; t < T; t++) if (l[t] == 0) bm[t >> 3] |= (uint8_t)(1u << (t & 7));
    if (mapB > 1) q += huffman_section_fmt(bm, mapB, q, x9);
    else { memcpy(q, bm, mapB); q += mapB; }                       /* mc:2104-2107: raw */
    free(bm);
    /* c stream */
    q += huffman_section_fmt(c, T, q, x9);                         /* mc:2178 */
    /* matches */
    uint32_t *mp = (uint32_t *)malloc(sizeof(uint32_t) * (T ? T : 1));
    uint32_t *ml = (uint32_t *)malloc(sizeof(uint32_t) * (T ? T : 1));
    uint32_t Pc = 0;
    for (uint32_t t = 0; t < T; t++) if (l[t] != 0) { mp[Pc] = p[t]; ml[Pc] = l[t]; Pc++; }
    put_u32(q, Pc); q += 4;                                        /* mc:2186-2188 */
    uint8_t *pk = (uint8_t *)malloc((size_t)(FCXO_P_BITS * Pc) / 8 + 1);
    uint32_t pkB = fcxo_combine_bits(mp, Pc, FCXO_P_BITS, pk);     /* mc:2193-2195 */
    q += huffman_section_fmt(pk, pkB, q, x9);                      /* mc:2207 */
    free(pk);
    uint32_t *gw = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)Pc * 3 + 1));
    uint32_t G = fcxo_golomb_encode(ml, Pc, gw);                   /* mc:2221 */
    put_u32(q, G); q += 4;                                         /* mc:2225-2227 */
    q += huffman_section_fmt((const uint8_t *)gw, G * 4, q, x9);   /* mc:2238 (emits nothing when G==0) */
    free(gw); free(mp); free(ml);
    return (uint32_t)(q - out);
}

uint32_t fcxo_serialize_block(const uint32_t *p, const uint32_t *l, const uint8_t *c, uint32_t T, uint8_t *out)
{
    return serialize_block_fmt(p, l, c, T, out, 0);
}

static uint32_t compress_block_fmt(const uint8_t *x, uint32_t n, uint8_t *out, int x9)
{
    if (!x || !out) return 0;                                      /* mc:2122-2123 */
    uint32_t *p = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t *l = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint8_t *c = (uint8_t *)malloc(n ? n : 1);
    uint32_t T = fcxo_lz77_parse(x, n, p, l, c);
    uint32_t bytes = serialize_block_fmt(p, l, c, T, out, x9);
    free(p); free(l); free(c);
    return bytes;
}

uint32_t fcxo_compress_block(const uint8_t *x, uint32_t n, uint8_t *out) { return compress_block_fmt(x, n, out, 0); }
uint32_t fcxo_compress_block_x9(const uint8_t *x, uint32_t n, uint8_t *out) { return compress_block_fmt(x, n, out, 1); }

uint64_t fcxo_compress_bound(uint64_t n)
{
    uint64_t nb = (n + FCXO_BLOCK_BYTES - 1) / FCXO_BLOCK_BYTES;
    /* per block: <= 9/8 n + 4 sections * (1+64+510+4+4) + 12 counters, + u32 size */
    return 10 + nb * (4 + 12 + 4 * 583 + 8) + n + n / 8 + 64;
}

static uint64_t compress_blocks_fmt(const uint8_t *x, uint64_t n, uint32_t b0, uint32_t b1, uint8_t *out, int x9)
{
    uint8_t *q = out;
    for (uint32_t b = b0; b < b1; b++) {
        uint64_t off = (uint64_t)b * FCXO_BLOCK_BYTES;
        uint32_t len = (uint32_t)((n - off) < FCXO_BLOCK_BYTES ? (n - off) : FCXO_BLOCK_BYTES);
        uint32_t sz = compress_block_fmt(x + off, len, q + 4, x9);
        put_u32(q, sz);                                            /* mc:4111 */
        q += 4 + sz;
    }
    return (uint64_t)(q - out);
}

uint64_t fcxo_compress_blocks(const uint8_t *x, uint64_t n, uint32_t b0, uint32_t b1, uint8_t *out)
{
    return compress_blocks_fmt(x, n, b0, b1, out, 0);
}

/* FCX9 container: 'F','C','X','9' | u64 bytes | u32 blocks | FCX9 records (include/fcx.h) */
uint64_t fcxo_compress_file_x9(const uint8_t *x, uint64_t n, uint8_t *out)
{
    uint64_t nb = (n + FCXO_BLOCK_BYTES - 1) / FCXO_BLOCK_BYTES;
    uint32_t nb32 = (uint32_t)nb;
    out[0] = 'F'; out[1] = 'C'; out[2] = 'X'; out[3] = '9';
    memcpy(out + 4, &n, 8);
    memcpy(out + 12, &nb32, 4);
    return 16 + compress_blocks_fmt(x, n, 0, nb32, out + 16, 1);
}

uint64_t fcxo_compress_file(const uint8_t *x, uint64_t n, uint8_t *out)
{
    uint64_t nb = (n + FCXO_BLOCK_BYTES - 1) / FCXO_BLOCK_BYTES;
    out[0] = 'F'; out[1] = 'C'; out[2] = 'X'; out[3] = '7';        /* mc:111,4078-4079 */
    put_u32(out + 4, (uint32_t)n);                                 /* u32 wrap, mc:105 */
    uint16_t nb16 = (uint16_t)nb;                                  /* u16 wrap, mc:106 */
    memcpy(out + 8, &nb16, 2);
    return 10 + fcxo_compress_blocks(x, n, 0, (uint32_t)nb, out + 10);
}

/* ------------------------------------------------------------------ */
/* decoder side                                                        */
/* ------------------------------------------------------------------ */
typedef struct { const uint8_t *p; uint64_t n; uint64_t pos; } rd_t;
/* the reference decodes from a zero-filled 2 MiB buffer (mc:4073,4195): reads past
 * the payload see zeros (App. B8). */
static uint8_t rd_u8(rd_t *r) { uint8_t v = r->pos < r->n ? r->p[r->pos] : 0; r->pos++; return v; }
static uint32_t rd_u32(rd_t *r)
{
    uint32_t v = 0;
    for (int k = 0; k < 4; k++) v |= (uint32_t)rd_u8(r) << (8 * k);
    return v;
}

/* mc:1107-1187 + mc:930-984.  Decodes cnt symbols into out (pre-zeroed by caller). */
static int huff_section_decode(rd_t *r, uint8_t *out, uint32_t cnt, int x9)
{
    uint32_t Tn = rd_u8(r);
    if (x9 && Tn == 0) {                             /* FCX9: the named symbol, then W (= 0) */
        uint8_t sym = rd_u8(r);
        uint32_t W0 = rd_u32(r);
        if (r->pos + (uint64_t)W0 * 4 > r->n + (2u << 20)) return -1;
        r->pos += (uint64_t)W0 * 4;
        memset(out, sym, cnt);
        return 0;
    }
    uint32_t mapB = (2 * Tn + 7) / 8;
    uint8_t map[64] = {0}, pair[510];
    for (uint32_t k = 0; k < mapB; k++) map[k] = rd_u8(r);
    for (uint32_t k = 0; k < 2 * Tn; k++) pair[k] = rd_u8(r);
    uint32_t W = rd_u32(r);
    if (r->pos + (uint64_t)W * 4 > r->n + (2u << 20)) return -1;
    uint32_t node = Tn - 1, j = 0;                   /* root = last node (mc:945-946); Tn==0 wraps */
    for (uint32_t wi = 0; wi < W; wi++) {
        uint32_t word = rd_u32(r);
        if (j >= cnt) continue;                      /* reference returned already; still consume */
        for (uint32_t b = 0; b < 32; b++) {
            if (node >= Tn) return -1;               /* reference would index out of bounds */
            uint32_t side = (word >> b) & 1;
            uint32_t ci = pair[2 * node + side];
            int internal = (map[(2 * node + side) >> 3] >> ((2 * node + side) & 7)) & 1;
            if (!internal) {                         /* mc:968: child index < 256 => leaf */
                out[j++] = (uint8_t)ci;
                node = Tn - 1;
                if (j >= cnt) break;                 /* mc:971-973 */
            } else {
                /* stored byte = treeIdx-256; simple index = treeIdx-256-(256-R), R = Tn+1 (mc:976-977) */
                node = ci - (256 - (Tn + 1));
            }
        }
    }
    return 0;
}

static uint32_t decompress_block_fmt(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t out_cap, int x9)
{
    rd_t r = { payload, nbytes, 0 };
    uint32_t T = rd_u32(&r);
    if (T > (2u << 20)) return UINT32_MAX;
    uint32_t mapB = (T + 7) / 8;
    uint8_t *bm = (uint8_t *)calloc(mapB + 1, 1);
    uint8_t *cs = (uint8_t *)calloc(T + 1, 1);
    uint32_t rc = UINT32_MAX;
    uint8_t *pk = NULL, *gb = NULL; uint32_t *ml = NULL;
    if (mapB > 1) { if (huff_section_decode(&r, bm, mapB, x9)) goto done; }
    else for (uint32_t k = 0; k < mapB; k++) bm[k] = rd_u8(&r);      /* mc:2277-2280 */
    if (huff_section_decode(&r, cs, T, x9)) goto done;               /* mc:2283 */
    uint32_t Pc = rd_u32(&r);
    if (Pc > (1u << 20)) goto done;
    uint32_t pkB = (FCXO_P_BITS * Pc) / 8 + 1;
    pk = (uint8_t *)calloc(pkB + 4, 1);
    if (huff_section_decode(&r, pk, pkB, x9)) goto done;             /* mc:2292 */
    uint32_t G = rd_u32(&r);
    if (G > (1u << 20)) goto done;
    gb = (uint8_t *)calloc((size_t)G * 4 + 4, 1);
    if (G || !x9) { if (huff_section_decode(&r, gb, G * 4, x9)) goto done; }   /* mc:2308 (FCX7, G==0: reads zeros, B8; FCX9: no section) */
    /* Golomb decode (mc:309-358): stops after Pc values */
    ml = (uint32_t *)calloc(Pc + 1, sizeof(uint32_t));
    {
        uint32_t got = 0, q = 0, rbits = 0, rv = 0; int after = 0;
        for (uint64_t pos = 0; pos < (uint64_t)G * 32 && got < Pc; pos++) {
            uint32_t bit = (gb[pos >> 3] >> (pos & 7)) & 1;
            if (!after) { if (bit) q++; else after = 1; }
            else {
                rv |= bit << rbits; rbits++;
                if (rbits >= 2) { ml[got++] = q * 4 + rv; q = 0; rv = 0; rbits = 0; after = 0; }
            }
        }
        /* a short Golomb stream leaves the reference indexing past its vector; flag it */
        uint32_t need = 0;
        for (uint32_t t = 0; t < T; t++) if (!((bm[t >> 3] >> (t & 7)) & 1)) need++;
        if (got < Pc && got < need) goto done;
    }
    {
        uint32_t cur = 0, m = 0;
        for (uint32_t t = 0; t < T; t++) {
            uint32_t l = 0, p = 0;
            if (!((bm[t >> 3] >> (t & 7)) & 1)) {
                if (m >= Pc) break;                                  /* mc:2337-2340 "Fatal Error": stop */
                /* decombine_bits (mc:1315-1338) */
                for (uint32_t k = 0; k < FCXO_P_BITS; k++) {
                    uint64_t bp = (uint64_t)m * FCXO_P_BITS + k;
                    p |= (uint32_t)((pk[bp >> 3] >> (bp & 7)) & 1) << k;
                }
                l = ml[m++];
            }
            if (p > cur && l > 0) goto done;                         /* reference would read before its buffer */
            if ((uint64_t)cur + l + 1 > out_cap) goto done;
            for (uint32_t k = 0; k < l; k++) out[cur + k] = out[cur - p + k]; /* mc:1730-1732 */
            cur += l;
            out[cur++] = cs[t];                                      /* mc:1734 */
        }
        rc = cur;
    }
done:
    free(bm); free(cs); free(pk); free(gb); free(ml);
    return rc;
}

uint32_t fcxo_decompress_block(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t out_cap)
{
    return decompress_block_fmt(payload, nbytes, out, out_cap, 0);
}
uint32_t fcxo_decompress_block_x9(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t out_cap)
{
    return decompress_block_fmt(payload, nbytes, out, out_cap, 1);
}

uint64_t fcxo_decompress_file(const uint8_t *file, uint64_t nbytes, uint8_t *out, uint64_t out_cap)
{
    if (nbytes < 10) return UINT64_MAX;
    if (file[0] != 'F' || file[1] != 'C' || file[2] != 'X') return UINT64_MAX;   /* mc:4146-4152 */
    if (file[3] != '7' && file[3] != '8' && file[3] != '9') return UINT64_MAX;    /* mc:4153-4164: '7' = lz77, '8' = lz78; '9' = opt-in */
    const int lz78 = file[3] == '8', x9 = file[3] == '9';
    uint32_t nb; uint64_t pos = 10, cur = 0;
    if (x9) {
        if (nbytes < 16) return UINT64_MAX;
        memcpy(&nb, file + 12, 4); pos = 16;
    } else { uint16_t nb16; memcpy(&nb16, file + 8, 2); nb = nb16; }
    for (uint32_t b = 0; b < nb; b++) {
        if (pos + 4 > nbytes) return UINT64_MAX;
        uint32_t sz = get_u32(file + pos); pos += 4;
        if (pos + sz > nbytes) return UINT64_MAX;
        uint64_t room = out_cap - cur;
        const uint32_t room32 = room > 0xffffffffu ? 0xffffffffu : (uint32_t)room;
        uint32_t got = lz78 ? fcxo_decompress_block_lz78(file + pos, sz, out + cur, room32)
                            : decompress_block_fmt(file + pos, sz, out + cur, room32, x9);
        if (got == UINT32_MAX) return UINT64_MAX;
        cur += got; pos += sz;
    }
    return cur;
}

/* ------------------------------------------------------------------ */
/* LZ78 / FCX8 block codec (scope row f-3, "next")      */
/* ------------------------------------------------------------------ */
/* Tokenizer, mc:1832-1899.  The reference keeps phrases in a map keyed by a BKDR hash and
 * verifies with memcmp (mc:1758-1796), so the hash never shows in the output: the token list
 * is plain LZ78 -- the longest known phrase, then one more byte, the new phrase taking the
 * token's 1-based number.  Here the dictionary is a trie held in an open-addressing table
 * keyed by (parent phrase, byte).  An input that ends inside a known phrase emits
 * (phrase, 0x00) and adds nothing (mc:1858-1862). */
uint32_t fcxo_lz78_tokenize(const uint8_t *x, uint32_t n, uint32_t *idx, uint8_t *c)
{
    if (n == 0) return 0;
    uint32_t cap = 1;
    while (cap < 2u * n + 16u) cap <<= 1;
    uint64_t *key = (uint64_t *)malloc((size_t)cap * sizeof(uint64_t));   /* (parent << 8 | byte) + 1, 0 = empty */
    uint32_t *val = (uint32_t *)malloc((size_t)cap * sizeof(uint32_t));
    memset(key, 0, (size_t)cap * sizeof(uint64_t));
    uint32_t T = 0, next = 1, cur = 0;
    while (cur < n) {
        uint32_t node = 0, k = 0;
        uint32_t slot = 0;
        while (cur + k < n) {
            const uint64_t kk = (((uint64_t)node << 8) | x[cur + k]) + 1;
            slot = (uint32_t)((kk * 0x9E3779B97F4A7C15ull) >> 40) & (cap - 1);
            while (key[slot] && key[slot] != kk) slot = (slot + 1) & (cap - 1);
            if (!key[slot]) break;                    /* phrase + this byte is new */
            node = val[slot];
            k++;
        }
        if (cur + k >= n) {                           /* ran out of input inside a known phrase (k >= 1) */
            idx[T] = node; c[T] = 0; T++;
            break;
        }
        key[slot] = (((uint64_t)node << 8) | x[cur + k]) + 1;
        val[slot] = next++;                           /* mc:1878: ++dictIdx after every emitted token */
        idx[T] = node; c[T] = x[cur + k]; T++;
        cur += k + 1;
    }
    free(key); free(val);
    return T;
}

/* create_huffman_tree (mc:535-617) for any number of leaves whose weights are ALL non-zero (the group
 * tree: every group of 256 ranks holds at least one used index).  child[2j], child[2j+1] = children of
 * internal node n + j; a value < n is a leaf.  Same two-queue rule as fcxo_huffman_tree. */
static void tree_general(const uint32_t *w, uint32_t n, uint32_t *child)
{
    uint32_t *ord = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    uint32_t *tmp = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    uint64_t *iw = (uint64_t *)malloc((size_t)n * sizeof(uint64_t));
    for (uint32_t i = 0; i < n; i++) ord[i] = i;
    /* stable bottom-up merge sort by weight (the reference's merge sort takes the left run on '<=') */
    for (uint32_t width = 1; width < n; width <<= 1) {
        for (uint32_t lo = 0; lo < n; lo += 2 * width) {
            uint32_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
            uint32_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) tmp[o++] = (w[ord[a]] <= w[ord[b]]) ? ord[a++] : ord[b++];
            while (a < mid) tmp[o++] = ord[a++];
            while (b < hi) tmp[o++] = ord[b++];
        }
        memcpy(ord, tmp, (size_t)n * sizeof(uint32_t));
    }
    uint32_t qa = 0, qb = 0, made = 0;
    while ((n - qa) + (made - qb) > 1) {
        uint64_t sum = 0;
        for (int k = 0; k < 2; k++) {
            int take_leaf;
            if (qa >= n) take_leaf = 0;
            else if (qb >= made) take_leaf = 1;
            else take_leaf = (uint64_t)w[ord[qa]] <= iw[qb];
            if (take_leaf) { child[2 * made + k] = ord[qa]; sum += w[ord[qa]]; qa++; }
            else { child[2 * made + k] = n + qb; sum += iw[qb]; qb++; }
        }
        iw[made++] = sum;
    }
    free(ord); free(tmp); free(iw);
}

/* my_compress_file_lz78 (mc:3127-3476): one FCX8 block payload; returns bytes written.
 *   u32 wCnt | u8 bitmap[maxIdx/8+1] | u32 groups | (groups-1) x {u32 left, u32 right} (groups > 1)
 *   | u32 T | u32 W, u32 words[W] (groups > 1) | u8 posInGroup[T] | HuffSection(c[T]) */
uint32_t fcxo_compress_block_lz78(const uint8_t *x, uint32_t n, uint8_t *out)
{
    if (n == 0) return 0;                              /* the reference is never called with an empty block */
    uint32_t *idx = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    uint8_t *c = (uint8_t *)malloc(n);
    const uint32_t T = fcxo_lz78_tokenize(x, n, idx, c);
    uint32_t maxIdx = 0;
    for (uint32_t t = 0; t < T; t++) if (idx[t] > maxIdx) maxIdx = idx[t];
    const uint32_t mapB = maxIdx / 8 + 1;              /* mc:3182-3183 */
    uint8_t *q = out;
    uint8_t *bm = q + 4;
    memset(bm, 0, mapB);
    for (uint32_t t = 0; t < T; t++) bm[idx[t] >> 3] |= (uint8_t)(1u << (idx[t] & 7));   /* mc:3189-3192 */
    /* rank of every used index among the used ones (the reference sorts the tokens and squeezes, mc:3159-3175) */
    uint32_t *rank = (uint32_t *)malloc(((size_t)maxIdx + 1) * sizeof(uint32_t));
    uint32_t wCnt = 0;
    for (uint32_t v = 0; v <= maxIdx; v++) { rank[v] = wCnt; wCnt += (bm[v >> 3] >> (v & 7)) & 1; }
    put_u32(q, wCnt); q += 4 + mapB;
    const uint32_t groups = (wCnt + 255) / 256;        /* mc:3227-3229 */
    put_u32(q, groups); q += 4;
    uint32_t *gchild = NULL;
    if (groups > 1) {
        uint32_t *gw = (uint32_t *)calloc(groups, sizeof(uint32_t));
        for (uint32_t t = 0; t < T; t++) gw[rank[idx[t]] >> 8]++;
        gchild = (uint32_t *)malloc((size_t)2 * (groups - 1) * sizeof(uint32_t));
        tree_general(gw, groups, gchild);
        free(gw);
        for (uint32_t j = 0; j < 2 * (groups - 1); j++) { put_u32(q, gchild[j]); q += 4; }   /* mc:3273-3287 */
    }
    put_u32(q, T); q += 4;                             /* mc:3321,3341 */
    if (groups > 1) {
        /* codes: root = last internal node, bit k of a code = k-th branch from the root, 0 = left (mc:2927-3006) */
        uint32_t nint = groups - 1;
        uint64_t *ibits = (uint64_t *)calloc(nint, sizeof(uint64_t));
        uint32_t *ilen = (uint32_t *)calloc(nint, sizeof(uint32_t));
        uint64_t *lbits = (uint64_t *)calloc(groups, sizeof(uint64_t));
        uint32_t *llen = (uint32_t *)calloc(groups, sizeof(uint32_t));
        int too_long = 0;
        for (int64_t j = (int64_t)nint - 1; j >= 0; j--)
            for (uint32_t side = 0; side < 2; side++) {
                const uint32_t ci = gchild[2 * j + side];
                const uint32_t d = ilen[j] + 1;
                if (d > 64) { too_long = 1; continue; }
                const uint64_t b = ibits[j] | ((uint64_t)side << ilen[j]);
                if (ci < groups) { lbits[ci] = b; llen[ci] = d; }
                else { ibits[ci - groups] = b; ilen[ci - groups] = d; }
            }
        uint64_t total = 0;
        for (uint32_t t = 0; t < T; t++) total += llen[rank[idx[t]] >> 8];
        const uint32_t W = (uint32_t)((total + 31) / 32);
        put_u32(q, W); q += 4;
        memset(q, 0, (size_t)W * 4);
        uint64_t pos = 0;
        for (uint32_t t = 0; t < T; t++) {
            const uint32_t g = rank[idx[t]] >> 8;
            for (uint32_t k = 0; k < llen[g]; k++, pos++)
                if ((lbits[g] >> k) & 1) q[pos >> 3] |= (uint8_t)(1u << (pos & 7));
        }
        q += (size_t)W * 4;
        free(ibits); free(ilen); free(lbits); free(llen); free(gchild);
        if (too_long) { free(idx); free(c); free(rank); return UINT32_MAX; }   /* > 64-bit group code: not representable here */
    }
    for (uint32_t t = 0; t < T; t++) *q++ = (uint8_t)(rank[idx[t]] & 255);   /* mc:3355 */
    q += fcxo_huffman_section(c, T, q);                /* mc:3384-3471 is the section builder of mc:1011-1091 again */
    free(idx); free(c); free(rank);
    return (uint32_t)(q - out);
}

/* my_decompress_file_lz78 (mc:3478-3710) + my_LZ78_decompress (mc:1901-1934); a trailing 0x00 is
 * always dropped (mc:3701-3703).  Returns decoded bytes or UINT32_MAX. */
uint32_t fcxo_decompress_block_lz78(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t out_cap)
{
    rd_t r = { payload, nbytes, 0 };
    const uint32_t wCnt = rd_u32(&r);
    if (wCnt == 0 || wCnt > (1u << 21)) return UINT32_MAX;
    uint32_t *used = (uint32_t *)malloc((size_t)wCnt * sizeof(uint32_t));
    uint32_t got = 0, v = 0;
    const uint64_t bm0 = r.pos;
    while (got < wCnt) {                                /* mc:3497-3503 */
        if (bm0 + (v >> 3) >= nbytes) { free(used); return UINT32_MAX; }
        if ((payload[bm0 + (v >> 3)] >> (v & 7)) & 1) used[got++] = v;
        v++;
    }
    r.pos += used[wCnt - 1] / 8 + 1;
    const uint32_t groups = rd_u32(&r);
    uint32_t *gchild = NULL;
    if (groups > 1) {
        if (groups > (1u << 13)) { free(used); return UINT32_MAX; }
        gchild = (uint32_t *)malloc((size_t)2 * (groups - 1) * sizeof(uint32_t));
        for (uint32_t j = 0; j < 2 * (groups - 1); j++) gchild[j] = rd_u32(&r);
    }
    const uint32_t T = rd_u32(&r);
    if (T == 0 || T > (1u << 21)) { free(used); free(gchild); return UINT32_MAX; }
    uint32_t *grp = (uint32_t *)calloc(T, sizeof(uint32_t));
    uint32_t rc = UINT32_MAX;
    uint8_t *cs = NULL; uint32_t *start = NULL, *len = NULL;
    if (groups > 1) {                                   /* mc:3008-3055 */
        const uint32_t W = rd_u32(&r);
        uint32_t node = groups - 2, j = 0;
        for (uint32_t wi = 0; wi < W; wi++) {
            const uint32_t word = rd_u32(&r);
            for (uint32_t b = 0; b < 32 && j < T; b++) {
                if (node >= groups - 1) goto done;
                const uint32_t ci = gchild[2 * node + ((word >> b) & 1)];
                if (ci < groups) { grp[j++] = ci; node = groups - 2; }
                else node = ci - groups;
            }
        }
    }
    {
        const uint64_t pos0 = r.pos;
        r.pos += T;
        cs = (uint8_t *)calloc(T, 1);
        if (huff_section_decode(&r, cs, T, 0) < 0) goto done;
        start = (uint32_t *)malloc((size_t)T * sizeof(uint32_t));
        len = (uint32_t *)malloc((size_t)T * sizeof(uint32_t));
        uint32_t cur = 0;
        for (uint32_t t = 0; t < T; t++) {
            const uint32_t pg = pos0 + t < nbytes ? payload[pos0 + t] : 0;
            const uint32_t ri = grp[t] * 256 + pg;
            if (ri >= wCnt) goto done;
            const uint32_t id = used[ri];
            uint32_t l = 0, s0 = 0;
            if (id) {                                   /* phrase id = token id-1 (mc:1921-1927) */
                if (id > t) goto done;
                l = len[id - 1]; s0 = start[id - 1];
            }
            if ((uint64_t)cur + l + 1 > out_cap) goto done;
            for (uint32_t k = 0; k < l; k++) out[cur + k] = out[s0 + k];
            out[cur + l] = cs[t];
            start[t] = cur; len[t] = l + 1;
            cur += l + 1;
        }
        if (cur && out[cur - 1] == 0) cur--;            /* mc:3701-3703 */
        rc = cur;
    }
done:
    free(used); free(gchild); free(grp); free(cs); free(start); free(len);
    return rc;
}

/* FCX8 container: the block loop of mc:4073-4136 with my_compress_file_lz78 per block (header byte '8',
 * mc:4080-4081).  out needs 10 + blocks * (4 + ~2.2 MiB). */
uint64_t fcxo_compress_file_lz78(const uint8_t *x, uint64_t n, uint8_t *out)
{
    uint64_t nb = (n + FCXO_BLOCK_BYTES - 1) / FCXO_BLOCK_BYTES;
    out[0] = 'F'; out[1] = 'C'; out[2] = 'X'; out[3] = '8';
    put_u32(out + 4, (uint32_t)n);
    uint16_t nb16 = (uint16_t)nb;
    memcpy(out + 8, &nb16, 2);
    uint8_t *q = out + 10;
    for (uint64_t b = 0; b < nb; b++) {
        uint64_t off = b * FCXO_BLOCK_BYTES;
        uint32_t len = (uint32_t)((n - off) < FCXO_BLOCK_BYTES ? (n - off) : FCXO_BLOCK_BYTES);
        uint32_t sz = fcxo_compress_block_lz78(x + off, len, q + 4);
        if (sz == UINT32_MAX) return UINT64_MAX;
        put_u32(q, sz);
        q += 4 + sz;
    }
    return (uint64_t)(q - out);
}

/* ------------------------------------------------------------------ */
/* mc:1407-1443 Sunday search, for the in-source known answers         */
/* ------------------------------------------------------------------ */
int32_t fcxo_sunday_search(const uint8_t *m, int32_t mlen, const uint8_t *s, int32_t slen)
{
    int32_t shift[256];
    for (int k = 0; k < 256; k++) shift[k] = slen + 1;   /* byte absent from the pattern */
    for (int32_t k = 0; k < slen; k++) shift[s[k]] = slen - k;
    int32_t at = 0;
    while (at + slen <= mlen) {
        int32_t k = 0;
        while (k < slen && m[at + k] == s[k]) k++;
        if (k == slen) return at;
        at += shift[m[at + slen]];
    }
    return -1;
}
