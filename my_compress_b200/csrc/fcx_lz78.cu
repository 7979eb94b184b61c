// fcx_lz78.cu -- the FCX8 (LZ78) block codec on the device (SURVEY 8f "next" row 3, App. E).
//
// Replaces my_LZ78_compress (mc:1832-1899) + my_compress_file_lz78 (mc:3127-3476, with
// huffman_encode_idxGroup mc:2927-3006) and, for decoding, my_decompress_file_lz78 (mc:3478-3710)
// + my_LZ78_decompress (mc:1901-1934).  Payload (all little-endian):
//   u32 wCnt | u8 idxBitmap[maxIdx/8+1] | u32 groups | (groups-1) x {u32 left, u32 right}   (groups > 1)
//   | u32 T | u32 W, u32 words[W] (groups > 1) | u8 posInGroup[T] | HuffSection(c[T])
//
// The tokenizer is plain LZ78: the reference's BKDR-hash map is verified with memcmp (mc:1758-1796), so the
// hash never shows in the output.  Every phrase depends on all earlier ones, so a block is one sequential
// walk; the parallelism is across blocks -- ONE THREAD PER BLOCK, 32 blocks per warp, every lane running the
// same per-byte step (probe the trie, descend or emit) on its own block.  The dictionary is a trie in an
// open-addressing table in HBM ((parent << 8 | byte) -> child, 2^20 eight-byte slots per block), with the 256
// children of the root in a direct table.  One dependent L2 / HBM probe per input byte: the kernel is
// latency-bound by construction and its throughput is the number of blocks in flight.
// Everything behind the tokens is parallel inside a block (one CTA per block): bitmap of the used ids,
// their ranks by prefix popcount, group weights, the group tree (same two-queue construction as the byte
// tree, leaves stably sorted first), group codes packed at scanned bit offsets, position bytes.  The
// character section is the lz77 path's Huffman stage run on the c stream (fcx_huff.cu, unchanged).
#include "fcx_internal.h"

#include <algorithm>

namespace fcx {

constexpr uint32_t L8_TABLE_SLOTS = 1u << 20;        // per block; T <= 256 + (n - 256) / 2 < slots / 2
constexpr uint32_t L8_MAX_TOKENS = 256 + (kBlockBytes - 256) / 2 + 16;
constexpr uint32_t L8_MAX_GROUPS = (L8_MAX_TOKENS + 255) / 256 + 1;      // 2050
constexpr uint32_t L8_STAGE_BYTES = 5u << 19;        // part A of a payload (everything before the char section): 2.5 MiB
constexpr uint32_t L8_MAP_WORDS = (L8_MAX_TOKENS + 31) / 32 + 1;

enum { L8M_T = 0, L8M_PARTA, L8M_ERR, L8M_WORDS = 4 };

struct Lz78Ws {
    unsigned long long *table;   // [blocks][L8_TABLE_SLOTS]  (key + 1) << 32 | child, 0 = empty
    uint32_t *idx;               // [blocks][L8_MAX_TOKENS]   phrase id per token, then its rank among the used ids
    uint32_t *wordRank;          // [blocks][L8_MAP_WORDS]    used ids below each bitmap word
    uint8_t *stage;              // [blocks][L8_STAGE_BYTES]  payload up to and including posInGroup
    uint32_t *meta;              // [blocks][L8M_WORDS]
    uint64_t *recOff;            // [blocks + 1] record offsets in the final container
    uint32_t maxBlocks;
};

// =========================================================================================
// tokens: one thread per block
// =========================================================================================
__device__ __forceinline__ uint32_t l8_hash(uint32_t kk)
{
    return (uint32_t)(((unsigned long long)kk * 0x9E3779B97F4A7C15ull) >> 44) & (L8_TABLE_SLOTS - 1);
}

__global__ void __launch_bounds__(32)
k78_tokenize(const uint8_t *__restrict__ in, const uint32_t *__restrict__ blkLen, Lz78Ws lw, uint8_t *__restrict__ cArena,
             uint32_t cStride, uint32_t nBlocks)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlocks) return;
    const uint32_t n = blkLen[b];
    const uint8_t *x = in + (size_t)b * kBlockBytes;
    unsigned long long *tab = lw.table + (size_t)b * L8_TABLE_SLOTS;
    uint32_t *idx = lw.idx + (size_t)b * L8_MAX_TOKENS;
    uint8_t *c = cArena + (size_t)b * cStride;
    uint32_t root[256];                                  // children of the empty phrase (local memory, L1-resident)
    for (int k = 0; k < 256; k++) root[k] = 0;
    // A flat step loop: every trip a lane does ONE dictionary probe (the dependent HBM access) or emits a one-byte
    // phrase.  With the descent as an inner loop the lanes of a warp -- 32 different blocks -- waited for the longest
    // phrase among them before any of them could start its next token (a block of random bytes, 2.5 bytes per phrase,
    // moved at the pace of a neighbour with 30-byte phrases).
    uint32_t T = 0, next = 1, cur = 0, k = 0, node = 0;       // phrase start, bytes of it matched so far, its dictionary node
    while (cur < n) {
        if (k == 0) {
            const uint32_t b0 = x[cur];
            node = root[b0];
            if (node == 0) {                             // a new one-byte phrase
                root[b0] = next++;                       // mc:1878: ++dictIdx after every emitted token
                idx[T] = 0; c[T] = (uint8_t)b0; T++;
                cur += 1;
                continue;
            }
            k = 1;
        }
        if (cur + k >= n) {                              // ran out of input inside a known phrase (mc:1858-1862)
            idx[T] = node; c[T] = 0; T++;
            break;
        }
        const uint32_t by = x[cur + k];
        const unsigned long long kk = (unsigned long long)((node << 8) | by) + 1;
        uint32_t slot = l8_hash((uint32_t)kk);
        unsigned long long e;
        while ((e = tab[slot]) != 0 && (e >> 32) != kk) slot = (slot + 1) & (L8_TABLE_SLOTS - 1);
        if (e == 0) {                                    // phrase + this byte is new
            tab[slot] = (kk << 32) | next++;
            idx[T] = node; c[T] = (uint8_t)by; T++;
            cur += k + 1;
            k = 0;
        } else {                                         // known: descend
            node = (uint32_t)e;
            k++;
        }
    }
    lw.meta[b * L8M_WORDS + L8M_T] = T;
}

// =========================================================================================
// everything between the tokens and the character section: one CTA per block
// =========================================================================================
constexpr int L8_THREADS = 256;

struct L8Smem {
    uint32_t gw[L8_MAX_GROUPS];          // group weights
    uint16_t order[L8_MAX_GROUPS];       // leaves sorted by (weight, group)
    uint32_t child[2 * L8_MAX_GROUPS];   // children of internal node j: a value < groups is a leaf, else groups + internal index
    uint16_t parLeaf[L8_MAX_GROUPS];     // leaf -> internal node j (side in parSideLeaf bit)
    uint16_t parInt[L8_MAX_GROUPS];      // internal -> parent internal (0xFFFF root)
    uint8_t sideLeaf[L8_MAX_GROUPS], sideInt[L8_MAX_GROUPS];
    uint8_t len[L8_MAX_GROUPS];          // code length per group
    unsigned long long code[L8_MAX_GROUPS];
    unsigned long long iw[L8_MAX_GROUPS];   // internal node weights
    unsigned long long warpSum[8];
    uint32_t red[8];
    uint32_t maxIdx, wCnt, err;
    unsigned long long totalBits;
};

__device__ __forceinline__ unsigned long long l8_cta_scan_excl(L8Smem &sm, unsigned long long v, unsigned long long &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();                                     // warpSum may still be read from an earlier scan
    if (lane == 31) sm.warpSum[warp] = inc;
    __syncthreads();
    unsigned long long base = 0; total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const unsigned long long s = sm.warpSum[w]; if (w < warp) base += s; total += s; }
    return base + inc - v;
}

__device__ __forceinline__ void l8_put32(uint8_t *p, uint32_t v) { *reinterpret_cast<uint32_t *>(p) = v; }   // 4-byte aligned

__global__ void __launch_bounds__(L8_THREADS)
k78_encode(Lz78Ws lw, const uint8_t *__restrict__ cArena, uint32_t cStride)
{
    extern __shared__ __align__(16) uint8_t l8_smem_raw[];
    L8Smem &sm = *reinterpret_cast<L8Smem *>(l8_smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    uint32_t *meta = lw.meta + (size_t)b * L8M_WORDS;
    const uint32_t T = meta[L8M_T];
    uint32_t *idx = lw.idx + (size_t)b * L8_MAX_TOKENS;
    uint32_t *wordRank = lw.wordRank + (size_t)b * L8_MAP_WORDS;
    uint8_t *pay = lw.stage + (size_t)b * L8_STAGE_BYTES;
    if (tid == 0) { sm.maxIdx = 0; sm.err = 0; sm.totalBits = 0; }
    __syncthreads();
    // ---- largest id -> bitmap size (mc:3182-3183) ----------------------------------------------------------
    {
        uint32_t m = 0;
        for (uint32_t t = tid; t < T; t += L8_THREADS) m = max(m, idx[t]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
        if ((tid & 31) == 0) atomicMax(&sm.maxIdx, m);
    }
    __syncthreads();
    const uint32_t maxIdx = sm.maxIdx;
    const uint32_t mapB = maxIdx / 8 + 1;
    const uint32_t mapW = (mapB + 3) / 4;
    uint32_t *bm = reinterpret_cast<uint32_t *>(pay + 4);            // the bitmap IS the payload's (word-aligned)
    for (uint32_t w = tid; w < mapW + 1; w += L8_THREADS) bm[w] = 0;  // + the word the next field starts in
    __syncthreads();
    for (uint32_t t = tid; t < T; t += L8_THREADS) { const uint32_t v = idx[t]; atomicOr(&bm[v >> 5], 1u << (v & 31)); }   // mc:3189-3192
    __syncthreads();
    // ---- rank of every used id: prefix popcount over the bitmap words ------------------------------------------
    {
        const uint32_t per = (mapW + L8_THREADS - 1) / L8_THREADS;
        const uint32_t w0 = min(mapW, tid * per), w1 = min(mapW, w0 + per);
        uint32_t s = 0;
        for (uint32_t w = w0; w < w1; w++) s += (uint32_t)__popc(bm[w]);
        unsigned long long total;
        uint32_t run = (uint32_t)l8_cta_scan_excl(sm, s, total);
        for (uint32_t w = w0; w < w1; w++) { wordRank[w] = run; run += (uint32_t)__popc(bm[w]); }
        if (tid == 0) sm.wCnt = (uint32_t)total;
    }
    __syncthreads();
    const uint32_t wCnt = sm.wCnt;
    const uint32_t groups = (wCnt + 255) / 256;                       // mc:3227-3229
    for (uint32_t g = tid; g < groups; g += L8_THREADS) sm.gw[g] = 0;
    __syncthreads();
    for (uint32_t t = tid; t < T; t += L8_THREADS) {                  // ids become ranks; group weights (mc:3264-3272)
        const uint32_t v = idx[t];
        const uint32_t r = wordRank[v >> 5] + (uint32_t)__popc(bm[v >> 5] & ((1u << (v & 31)) - 1u));
        idx[t] = r;
        atomicAdd(&sm.gw[r >> 8], 1u);
    }
    __syncthreads();
    uint8_t *q = pay;
    if (tid == 0) l8_put32(q, wCnt);
    q += 4 + mapB;                                                    // from here on fields are not word-aligned
    auto put32u = [](uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); };
    if (tid == 0) put32u(q, groups);
    q += 4;
    if (groups > 1) {
        // ---- group tree: stable sort by weight, two-queue merge (leaf first on ties, older internal node first) ---------
        for (uint32_t g = tid; g < groups; g += L8_THREADS) {
            const uint32_t w = sm.gw[g];
            uint32_t rank = 0;
            for (uint32_t h = 0; h < groups; h++) { const uint32_t v = sm.gw[h]; rank += (v < w || (v == w && h < g)) ? 1u : 0u; }
            sm.order[rank] = (uint16_t)g;
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t n = groups;
            uint32_t qa = 0, qb = 0, made = 0;
            while ((n - qa) + (made - qb) > 1) {
                unsigned long long sum = 0;
                for (int k = 0; k < 2; k++) {
                    bool leaf;
                    if (qa >= n) leaf = false;
                    else if (qb >= made) leaf = true;
                    else leaf = (unsigned long long)sm.gw[sm.order[qa]] <= sm.iw[qb];
                    if (leaf) {
                        const uint32_t g = sm.order[qa];
                        sm.child[2 * made + k] = g; sm.parLeaf[g] = (uint16_t)made; sm.sideLeaf[g] = (uint8_t)k;
                        sum += sm.gw[g]; qa++;
                    } else {
                        sm.child[2 * made + k] = n + qb; sm.parInt[qb] = (uint16_t)made; sm.sideInt[qb] = (uint8_t)k;
                        sum += sm.iw[qb]; qb++;
                    }
                }
                sm.iw[made++] = sum;
            }
            sm.parInt[n - 2] = 0xFFFF;                                // root = last internal node
        }
        __syncthreads();
        for (uint32_t j = tid; j < 2 * (groups - 1); j += L8_THREADS) put32u(q + 4 * j, sm.child[j]);   // mc:3273-3287
        // codes: every leaf climbs to the root; bit k of a code = k-th branch from the root, 0 = left (mc:2927-3006)
        for (uint32_t g = tid; g < groups; g += L8_THREADS) {
            unsigned long long code = sm.sideLeaf[g];
            uint32_t len = 1, p = sm.parLeaf[g];
            while (sm.parInt[p] != 0xFFFF && len < 65) { code = (code << 1) | sm.sideInt[p]; len++; p = sm.parInt[p]; }
            if (len > 64) { sm.err = 1; len = 64; }
            sm.code[g] = code; sm.len[g] = (uint8_t)len;
        }
        q += 8 * (groups - 1);
    }
    __syncthreads();
    if (tid == 0) put32u(q, T);                                       // mc:3321,3341
    q += 4;
    if (groups > 1) {
        // ---- group codes of all tokens at scanned bit offsets ----------------------------------------------------------
        const uint32_t per = (T + L8_THREADS - 1) / L8_THREADS;
        const uint32_t t0 = min(T, tid * per), t1 = min(T, t0 + per);
        unsigned long long bits = 0;
        for (uint32_t t = t0; t < t1; t++) bits += sm.len[idx[t] >> 8];
        unsigned long long total;
        unsigned long long pos = l8_cta_scan_excl(sm, bits, total);
        const uint32_t W = (uint32_t)((total + 31) / 32);
        if (tid == 0) put32u(q, W);
        q += 4;
        // the words sit at an arbitrary byte offset: assemble in an aligned scratch view (wordRank is free now: W <= T * 2 words
        // does not fit it in general, so write bytes through 32-bit atomics on the aligned words around q)
        uint8_t *wb = q;
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(wb) & 3);
        uint32_t *aw = reinterpret_cast<uint32_t *>(wb - mis);       // aligned words covering [wb, wb + 4W)
        const unsigned long long shiftBits = 8ull * mis;
        const unsigned long long nAw = (shiftBits + 32ull * W + 31) / 32;
        // zero what the codes will be OR-ed into, keeping the bytes before wb (the W field / earlier fields)
        __syncthreads();
        for (unsigned long long w = tid; w < nAw; w += L8_THREADS) {
            if (w == 0 && mis) aw[0] &= (1u << shiftBits) - 1u; else aw[w] = 0;
        }
        __syncthreads();
        {
            unsigned long long p = pos + shiftBits;
            for (uint32_t t = t0; t < t1; t++) {
                const uint32_t g = idx[t] >> 8;
                const unsigned long long cde = sm.code[g];
                const uint32_t len = sm.len[g];
                // up to 64 bits at bit offset p: three 32-bit pieces
                const uint32_t sh = (uint32_t)(p & 31);
                const unsigned long long w0 = p >> 5;
                const uint32_t lo = (uint32_t)cde << sh;
                const unsigned long long mid = sh ? (cde >> (32 - sh)) : (cde >> 32 << 0);
                if (lo) atomicOr(&aw[w0], lo);
                if (sh) {
                    const uint32_t m1 = (uint32_t)mid, m2 = (uint32_t)(mid >> 32);
                    if (m1) atomicOr(&aw[w0 + 1], m1);
                    if (m2) atomicOr(&aw[w0 + 2], m2);
                } else {
                    const uint32_t m1 = (uint32_t)(cde >> 32);
                    if (m1) atomicOr(&aw[w0 + 1], m1);
                }
                p += len;
            }
        }
        q += 4ull * W;
        __syncthreads();
    }
    for (uint32_t t = tid; t < T; t += L8_THREADS) q[t] = (uint8_t)(idx[t] & 255);      // mc:3355
    q += T;
    if (tid == 0) {
        meta[L8M_PARTA] = (uint32_t)(q - pay);
        meta[L8M_ERR] = sm.err | ((q - pay) > (ptrdiff_t)(L8_STAGE_BYTES - 8) ? 2u : 0u);
    }
}

// meta words of the lz77 pipeline's Huffman stage: only the c stream exists (T symbols)
__global__ void k78_huff_meta(Ws ws, Lz78Ws lw, uint32_t nBlocks)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlocks) return;
    const uint32_t T = lw.meta[b * L8M_WORDS + L8M_T];
    uint32_t *m = ws.meta + (size_t)b * M_WORDS;
    m[M_T] = T; m[M_PC] = 0; m[M_G] = 0; m[M_GBITS] = 0;
    m[M_S0] = 0; m[M_S1] = T; m[M_S2] = 0; m[M_S3] = 0;
}

// record offsets in the final container: u32 size | part A | char section
__global__ void __launch_bounds__(1024)
k78_layout(Ws ws, Lz78Ws lw, uint32_t nBlocks, uint32_t withHeader, uint64_t outCap, unsigned long long *cursor, uint32_t *errFlag)
{
    __shared__ unsigned long long warpSum[32];
    __shared__ unsigned long long running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = *cursor + (withHeader ? 10ull : 0ull);
    __syncthreads();
    for (uint32_t base = 0; base < nBlocks; base += 1024) {
        const uint32_t b = base + tid;
        unsigned long long rec = 0;
        if (b < nBlocks) {
            const uint32_t *si = ws.segInfo + ((size_t)b * 4 + 1) * SI_WORDS;
            const uint32_t sec = si[SI_S] ? si[SI_HDR] + 4 + 4 * si[SI_W] : 0;
            rec = 4ull + lw.meta[b * L8M_WORDS + L8M_PARTA] + sec;
            if (lw.meta[b * L8M_WORDS + L8M_ERR]) atomicOr(errFlag, (uint32_t)E_CODELEN);
        }
        unsigned long long inc = rec;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warpSum[warp] = inc;
        __syncthreads();
        unsigned long long wbase = 0, total = 0;
        for (int w = 0; w < 32; w++) { const unsigned long long v = warpSum[w]; if (w < warp) wbase += v; total += v; }
        if (b < nBlocks) lw.recOff[b] = running + wbase + inc - rec;
        __syncthreads();
        if (tid == 0) running += total;
        __syncthreads();
    }
    if (tid == 0) {
        lw.recOff[nBlocks] = running;
        *cursor = running;
        if (running > outCap) atomicOr(errFlag, (uint32_t)E_CAPACITY);
    }
}

// copies part A and the char section of every block behind its u32 size; block 0 of the file writes the header
__global__ void __launch_bounds__(256)
k78_assemble(Ws ws, Lz78Ws lw, const uint8_t *__restrict__ tmp, uint8_t *__restrict__ out, uint32_t withHeader,
             uint32_t hdrTotalBytes, uint32_t hdrBlocks, unsigned long long batchStart, const uint32_t *errFlag)
{
    if (*errFlag) return;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (b == 0 && withHeader && tid == 0) {
        uint8_t *h = out + batchStart;                    // mc:4078-4086 with '8' (mc:4080-4081)
        h[0] = 'F'; h[1] = 'C'; h[2] = 'X'; h[3] = '8';
        st_u32_bytes(h + 4, hdrTotalBytes);
        h[8] = (uint8_t)hdrBlocks; h[9] = (uint8_t)(hdrBlocks >> 8);
    }
    const unsigned long long off = lw.recOff[b];
    const uint32_t partA = lw.meta[b * L8M_WORDS + L8M_PARTA];
    const uint32_t size = (uint32_t)(lw.recOff[b + 1] - off - 4);
    uint8_t *q = out + off;
    if (tid == 0) st_u32_bytes(q, size);
    const uint8_t *a = lw.stage + (size_t)b * L8_STAGE_BYTES;
    for (uint32_t i = tid; i < partA; i += 256) q[4 + i] = a[i];
    // the Huffman stage wrote "u32 size | u32 T | cSec | u32 Pc | u32 G" records of its own into tmp
    const uint8_t *s = tmp + ws.blkOff[b] + 8;
    const uint32_t sec = size - partA;
    for (uint32_t i = tid; i < sec; i += 256) q[4 + partA + i] = s[i];
}

// =========================================================================================
// host side
// =========================================================================================
struct Lz78Host {
    Lz78Ws lw{};
    uint8_t *tmp = nullptr;          // the Huffman stage's own container for the batch
    uint64_t tmpCap = 0;
    unsigned long long *cursor = nullptr;   // device: running output cursor + error flag
    std::vector<void *> allocs;
};

static Lz78Host *lz78_get(fcx_ctx *c, uint32_t blocks, std::string &err)
{
    Lz78Host *h = reinterpret_cast<Lz78Host *>(c->lz78);
    if (h && h->lw.maxBlocks >= blocks) return h;
    if (h) { for (void *p : h->allocs) cudaFree(p); delete h; c->lz78 = nullptr; }
    h = new Lz78Host();
    auto A = [&](void **p, size_t bytes) {
        if (cudaMalloc(p, bytes + 256) != cudaSuccess) { err = "lz78 workspace allocation failed"; return false; }
        h->allocs.push_back(*p);
        return true;
    };
    const size_t B = blocks;
    bool ok = A(reinterpret_cast<void **>(&h->lw.table), B * L8_TABLE_SLOTS * 8) &&
              A(reinterpret_cast<void **>(&h->lw.idx), B * L8_MAX_TOKENS * 4) &&
              A(reinterpret_cast<void **>(&h->lw.wordRank), B * L8_MAP_WORDS * 4) &&
              A(reinterpret_cast<void **>(&h->lw.stage), B * (size_t)L8_STAGE_BYTES) &&
              A(reinterpret_cast<void **>(&h->lw.meta), B * L8M_WORDS * 4) &&
              A(reinterpret_cast<void **>(&h->lw.recOff), (B + 1) * 8) &&
              A(reinterpret_cast<void **>(&h->cursor), 64);
    h->tmpCap = fcx_compress_bound(B * kBlockBytes);
    ok = ok && A(reinterpret_cast<void **>(&h->tmp), h->tmpCap);
    if (!ok) { for (void *p : h->allocs) cudaFree(p); delete h; cudaGetLastError(); return nullptr; }
    h->lw.maxBlocks = blocks;
    cudaFuncSetAttribute(k78_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(L8Smem));
    c->lz78 = h;
    return h;
}

void lz78_release(fcx_ctx *c)
{
    Lz78Host *h = reinterpret_cast<Lz78Host *>(c->lz78);
    if (!h) return;
    for (void *p : h->allocs) cudaFree(p);
    delete h;
    c->lz78 = nullptr;
}

__global__ void k78_set_blklen(uint32_t *blkLen, uint64_t n, uint32_t nb)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint64_t off = (uint64_t)b * kBlockBytes;
    const uint64_t rem = n > off ? n - off : 0;
    blkLen[b] = (uint32_t)(rem < kBlockBytes ? rem : kBlockBytes);
}

// Compresses n bytes at d_in (device) into FCX8 records (+ header) at d_out; *outBytes = bytes written.
int lz78_compress_device(fcx_ctx *c, const uint8_t *d_in, uint64_t n, uint8_t *d_out, uint64_t outCap, bool withHeader,
                         uint64_t hdrBytes, uint32_t hdrBlocks, uint64_t *outBytes)
{
    const uint64_t nBlocks = (n + kBlockBytes - 1) / kBlockBytes;
    // blocks in flight are the tokenizer's only parallelism (one thread per block, one dependent probe per byte):
    // a batch is as large as the context allows -- 12.6 MiB of lz78 workspace per block, 13 GB for 1024 blocks
    const uint32_t batch = std::min<uint32_t>(c->maxBlocks, kMaxBatchBlocks);
    std::string err;
    Lz78Host *h = lz78_get(c, (uint32_t)std::min<uint64_t>(std::max<uint64_t>(nBlocks, 1), batch), err);
    if (!h) return fail(FCX_ERR_CUDA, err);
    cudaStream_t st = c->stream;
    const Ws &w = c->ws;
    CK(cudaMemsetAsync(h->cursor, 0, 64, st));
    uint32_t *errFlag = reinterpret_cast<uint32_t *>(h->cursor + 1);
    if (nBlocks == 0) {
        if (withHeader) {
            uint8_t hd[10] = {'F', 'C', 'X', '8', 0, 0, 0, 0, 0, 0};
            const uint32_t tb = (uint32_t)hdrBytes; const uint16_t bn = (uint16_t)hdrBlocks;
            memcpy(hd + 4, &tb, 4); memcpy(hd + 8, &bn, 2);
            if (outCap < 10) return fail(FCX_ERR_CAPACITY, "output buffer too small");
            CK(cudaMemcpyAsync(d_out, hd, 10, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
        }
        *outBytes = withHeader ? 10 : 0;
        return FCX_OK;
    }
    for (uint64_t first = 0; first < nBlocks; first += batch) {
        const uint32_t nb = (uint32_t)std::min<uint64_t>(batch, nBlocks - first);
        const uint8_t *bin = d_in + first * kBlockBytes;
        const uint64_t rem = n - first * kBlockBytes;
        CK(cudaMemsetAsync(h->lw.table, 0, (size_t)nb * L8_TABLE_SLOTS * 8, st));
        k78_set_blklen<<<(nb + 255) / 256, 256, 0, st>>>(w.blkLen, rem, nb);
        k78_tokenize<<<(nb + 31) / 32, 32, 0, st>>>(bin, w.blkLen, h->lw, w.strm[1], w.strmStride[1], nb);
        k78_encode<<<nb, L8_THREADS, sizeof(L8Smem), st>>>(h->lw, w.strm[1], w.strmStride[1]);
        // character sections: the lz77 path's Huffman stage on the c streams, into a scratch container
        k78_huff_meta<<<(nb + 255) / 256, 256, 0, st>>>(w, h->lw, nb);
        CK(cudaMemsetAsync(h->tmp, 0, std::min<uint64_t>(h->tmpCap, fcx_compress_bound((uint64_t)nb * kBlockBytes)), st));
        CK(cudaMemsetAsync(w.counters, 0, C_WORDS * sizeof(uint64_t) + TOT_WORDS * sizeof(uint32_t), st));
        c->launches += 6 + launch_huff_stage(w, nb, h->tmp, h->tmpCap, 0, 0, 0, 0, c->numSM, st, nullptr, nullptr);
        const bool hdr = withHeader && first == 0;
        k78_layout<<<1, 1024, 0, st>>>(w, h->lw, nb, hdr ? 1u : 0u, outCap, h->cursor, errFlag);
        k78_assemble<<<nb, 256, 0, st>>>(w, h->lw, h->tmp, d_out, hdr ? 1u : 0u, (uint32_t)hdrBytes, hdrBlocks, 0, errFlag);
        CK(cudaGetLastError());
    }
    unsigned long long res[2];
    CK(cudaMemcpyAsync(res, h->cursor, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t e = (uint32_t)res[1];
    if (e & E_CAPACITY) return fail(FCX_ERR_CAPACITY, "output buffer smaller than the compressed result");
    if (e) return fail(FCX_ERR_UNSUPPORTED, "a group code longer than 64 bits or an oversized payload (the reference overflows its buffers there)");
    *outBytes = res[0];
    return FCX_OK;
}

} // namespace fcx

// =========================================================================================
// FCX8 decoder: my_decompress_file_lz78 (mc:3478-3710) + my_LZ78_decompress (mc:1901-1934)
// =========================================================================================
// One CTA per block.  The used-id list comes from the bitmap by prefix popcount; the group codes and the character
// section are prefix-coded streams and go through the speculative decoder of fcx_bits.cuh; a token's phrase is its
// parent's phrase plus one byte, so phrase lengths are depths in a forest (pointer jumping over (ancestor, distance)
// pairs), output offsets are a scan of the lengths, and every output byte either is a token's own character or
// copies the byte at the same offset of the parent phrase -- the same pointer-doubling collapse as the FCX7 replay,
// over the whole block in HBM.  A trailing 0x00 is dropped unconditionally, as the reference does (mc:3701-3703).
#include "fcx_bits.cuh"

namespace fcx {

constexpr uint32_t D8_BAD = 0xFFFFFFFFu;
// scratch per block (bytes)
constexpr size_t D8_OFF_USED = 0;                                         // u32[L8_MAX_TOKENS + 64]  id of every rank
constexpr size_t D8_OFF_GRP = D8_OFF_USED + 4 * (size_t)(L8_MAX_TOKENS + 64);     // u16[T] group of every token
constexpr size_t D8_OFF_CS = D8_OFF_GRP + 2 * (size_t)(L8_MAX_TOKENS + 64);       // u8[T] characters (zero-filled first: App. B1)
constexpr size_t D8_OFF_ANC = D8_OFF_CS + (size_t)(L8_MAX_TOKENS + 64);           // u64[T] ancestor << 32 | distance
constexpr size_t D8_OFF_START = D8_OFF_ANC + 8 * (size_t)(L8_MAX_TOKENS + 64);    // u32[T + 1] output offset of every phrase
constexpr size_t D8_OFF_SRC = D8_OFF_START + 4 * (size_t)(L8_MAX_TOKENS + 64);    // u32[1 MiB + 64] source of every output byte
constexpr size_t D8_SCRATCH = ((D8_OFF_SRC + 4 * (size_t)(kBlockBytes + 64)) + 255) & ~(size_t)255;
static_assert(D8_OFF_ANC % 8 == 0, "64-bit pairs are aligned");
constexpr uint32_t D8_NONE = 0xFFFFFFFFu;

size_t lz78_decode_scratch_per_block() { return D8_SCRATCH; }

struct GroupCodec {                       // tree walk, one bit per step (mc:3008-3055)
    const uint32_t *child;                // [2 * (groups - 1)]
    uint32_t groups;
    template <class B>
    __device__ __forceinline__ int next(B &lb, uint32_t &sym) const
    {
        uint32_t node = groups - 2;       // root = last internal node
        for (;;) {
            if (lb.nbits == 0) { lb.fill(); if (lb.nbits == 0) return 1; }
            const uint32_t ci = child[2 * node + ((uint32_t)lb.buf & 1u)];
            lb.skip(1);
            if (ci < groups) { sym = ci; return 0; }
            node = ci - groups;
            if (node >= groups - 1) return 2;
        }
    }
};

struct D8Smem {
    SpecSmem spec;
    uint16_t lut[DEC_LUT];
    uint8_t pair[512];
    uint8_t map[64];
    uint32_t child[2 * L8_MAX_GROUPS];
    unsigned long long warpSum[DEC_WARPS];
    uint32_t bad, lastId, found;
};

__device__ __forceinline__ unsigned long long d8_cta_scan_excl(D8Smem &sm, unsigned long long v, unsigned long long &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();
    if (lane == 31) sm.warpSum[warp] = inc;
    __syncthreads();
    unsigned long long base = 0; total = 0;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; w++) { const unsigned long long s = sm.warpSum[w]; if (w < warp) base += s; total += s; }
    return base + inc - v;
}

__global__ void __launch_bounds__(DEC_THREADS)
k78_decode(const uint8_t *__restrict__ file, const uint64_t *__restrict__ recOff, uint8_t *__restrict__ out,
           uint32_t *__restrict__ outLen, uint8_t *__restrict__ scratchAll, size_t scratchPerBlock, size_t outStride,
           uint32_t outCapPerBlock)
{
    extern __shared__ __align__(16) uint8_t d8_smem_raw[];
    D8Smem &sm = *reinterpret_cast<D8Smem *>(d8_smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    const uint8_t *rec = file + recOff[b];
    const uint32_t plen = rd32(rec);
    const uint8_t *pl = rec + 4;
    uint8_t *scr = scratchAll + (size_t)b * scratchPerBlock;
    uint32_t *used = reinterpret_cast<uint32_t *>(scr + D8_OFF_USED);
    uint16_t *grp = reinterpret_cast<uint16_t *>(scr + D8_OFF_GRP);
    uint8_t *cs = scr + D8_OFF_CS;
    unsigned long long *anc = reinterpret_cast<unsigned long long *>(scr + D8_OFF_ANC);
    uint32_t *start = reinterpret_cast<uint32_t *>(scr + D8_OFF_START);
    uint32_t *src = reinterpret_cast<uint32_t *>(scr + D8_OFF_SRC);
    uint8_t *o = out + (size_t)b * outStride;                      // a block can decode to one byte more than 1 MiB before the 0x00 drop
    if (tid == 0) { sm.bad = 0; sm.lastId = 0; sm.found = 0; }
    __syncthreads();
    auto fail_block = [&]() { if (tid == 0) outLen[b] = D8_BAD; };
    if (plen < 16) { fail_block(); return; }
    const uint32_t wCnt = rd32(pl);
    if (wCnt == 0 || wCnt > L8_MAX_TOKENS) { fail_block(); return; }
    // ---- the used ids: the bitmap ends with the byte that holds the wCnt-th set bit (mc:3497-3503) -------------
    const uint8_t *bmB = pl + 4;
    const uint32_t maxBytes = min(plen - 4, L8_MAX_TOKENS / 8 + 8);
    {
        const uint32_t per = (maxBytes + DEC_THREADS - 1) / DEC_THREADS;
        const uint32_t y0 = min(maxBytes, tid * per), y1 = min(maxBytes, y0 + per);
        uint32_t s = 0;
        for (uint32_t y = y0; y < y1; y++) s += (uint32_t)__popc((uint32_t)bmB[y]);
        unsigned long long total;
        uint32_t run = (uint32_t)d8_cta_scan_excl(sm, s, total);
        for (uint32_t y = y0; y < y1 && run < wCnt; y++) {
            uint32_t m = bmB[y];
            while (m && run < wCnt) {
                const uint32_t bit = (uint32_t)__ffs(m) - 1;
                m &= m - 1;
                const uint32_t id = 8 * y + bit;
                used[run] = id;
                if (run == wCnt - 1) { sm.lastId = id; sm.found = 1; }
                run++;
            }
        }
    }
    __syncthreads();
    if (!sm.found) { fail_block(); return; }
    uint32_t at = 4 + sm.lastId / 8 + 1;
    if (at + 4 > plen) { fail_block(); return; }
    const uint32_t groups = rd32(pl + at); at += 4;
    if (groups == 0 || groups > L8_MAX_GROUPS) { fail_block(); return; }
    if (groups > 1) {
        if ((unsigned long long)at + 8ull * (groups - 1) + 4 > plen) { fail_block(); return; }
        for (uint32_t j = tid; j < 2 * (groups - 1); j += DEC_THREADS) sm.child[j] = rd32(pl + at + 4 * j);
        at += 8 * (groups - 1);
    }
    if (at + 4 > plen) { fail_block(); return; }
    const uint32_t T = rd32(pl + at); at += 4;
    if (T == 0 || T > L8_MAX_TOKENS) { fail_block(); return; }
    __syncthreads();
    // ---- group of every token --------------------------------------------------------------------------------------------
    if (groups > 1) {
        if (at + 4 > plen) { fail_block(); return; }
        const uint32_t W = rd32(pl + at); at += 4;
        if ((unsigned long long)at + 4ull * W > plen) { fail_block(); return; }
        GroupCodec gc{sm.child, groups};
        LaneBits<true> lb;
        lb.open(pl + at, W);
        bool gbad = false;
        const uint32_t got = spec_decode(sm.spec, gc, lb, T, [grp](uint32_t i, uint32_t sym) { grp[i] = (uint16_t)sym; }, gbad);
        // tokens the stream did not reach keep group 0 (the reference's calloc'ed array); a broken tree ends the block
        if (__syncthreads_or(gbad)) { fail_block(); return; }
        (void)got;
        at += 4 * W;
    }
    const uint32_t posAt = at;                                        // u8 posInGroup[T]
    if ((unsigned long long)at + T > plen) { fail_block(); return; }
    at += T;
    // ---- characters: one Huffman section (same layout as the lz77 path's, mc:3384-3471) --------------------------------
    {
        if (at + 5 > plen) { fail_block(); return; }
        const uint32_t Tn = pl[at];
        const uint32_t mapB = (2 * Tn + 7) / 8;
        const uint32_t wAt = at + 1 + mapB + 2 * Tn;
        if (wAt + 4 > plen) { fail_block(); return; }
        const uint32_t W = rd32(pl + wAt);
        if ((unsigned long long)wAt + 4 + 4ull * W > plen) { fail_block(); return; }
        if (Tn && W) {                                                // Tn == 0: single-symbol stream decodes to zeros (App. B1)
            for (uint32_t i = tid; i < mapB; i += DEC_THREADS) sm.map[i] = pl[at + 1 + i];
            for (uint32_t i = tid; i < 2 * Tn; i += DEC_THREADS) sm.pair[i] = pl[at + 1 + mapB + i];
            __syncthreads();
            build_huff_lut(sm.lut, sm.pair, sm.map, Tn);
            __syncthreads();
            HuffCodec cd{sm.lut, sm.pair, sm.map, Tn, Tn + 1, Tn - 1};
            LaneBits<true> lb;
            lb.open(pl + wAt + 4, W);
            bool cbad = false;
            spec_decode(sm.spec, cd, lb, T, [cs](uint32_t i, uint32_t sym) { cs[i] = (uint8_t)sym; }, cbad);
            if (__syncthreads_or(cbad)) { fail_block(); return; }
        }
    }
    __syncthreads();
    // ---- parents: token t extends phrase id (0: nothing); phrase k is token k - 1 (mc:1921-1927) ----------------------
    for (uint32_t t = tid; t < T; t += DEC_THREADS) {
        const uint32_t ri = (uint32_t)grp[t] * 256u + pl[posAt + t];
        uint32_t id = 0;
        if (ri >= wCnt) sm.bad = 1; else id = used[ri];
        if (id > t) { sm.bad = 1; id = 0; }
        anc[t] = ((unsigned long long)(id ? id - 1 : D8_NONE) << 32) | 1ull;      // (ancestor, bytes gathered so far)
    }
    __syncthreads();
    if (sm.bad) { fail_block(); return; }
    // parent of every token for the copy step, before the jumps overwrite the ancestors
    for (uint32_t t = tid; t < T; t += DEC_THREADS) start[t] = (uint32_t)(anc[t] >> 32);     // start[] doubles as parent[] for now
    __syncthreads();
    // ---- phrase lengths: pointer jumping; a pair read is one 64-bit load, so the in-place update is safe ----------------
    for (;;) {
        bool changed = false;
        for (uint32_t t = tid; t < T; t += DEC_THREADS) {
            const unsigned long long v = anc[t];
            const uint32_t a = (uint32_t)(v >> 32);
            if (a != D8_NONE) {
                const unsigned long long u = *reinterpret_cast<volatile unsigned long long *>(anc + a);
                anc[t] = (u & 0xFFFFFFFF00000000ull) | (uint32_t)((uint32_t)v + (uint32_t)u);
                changed = true;
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    // ---- output offsets: scan of the lengths; src[] of every byte ------------------------------------------------------------
    // (grp[] is free now: the per-token parent moves there?  no -- parents live in start[], lengths in anc[]; offsets go to src's tail)
    uint32_t total32;
    {
        const uint32_t per = (T + DEC_THREADS - 1) / DEC_THREADS;
        const uint32_t t0 = min(T, tid * per), t1 = min(T, t0 + per);
        unsigned long long s = 0;
        for (uint32_t t = t0; t < t1; t++) s += (uint32_t)anc[t];
        unsigned long long total;
        unsigned long long run = d8_cta_scan_excl(sm, s, total);
        if (total > outCapPerBlock) { fail_block(); return; }            // CTA-uniform
        total32 = (uint32_t)total;
        // offsets replace the lengths' upper halves: anc[t] = offset << 32 | length
        for (uint32_t t = t0; t < t1; t++) { const uint32_t l = (uint32_t)anc[t]; anc[t] = (run << 32) | l; run += l; }
    }
    __syncthreads();
    // every byte: the last byte of a phrase is the token's character, byte j < len - 1 copies byte j of the parent phrase
    for (uint32_t t = tid; t < T; t += DEC_THREADS) {
        const unsigned long long v = anc[t];
        const uint32_t off = (uint32_t)(v >> 32), l = (uint32_t)v;
        const uint32_t par = start[t];
        src[off + l - 1] = off + l - 1;
        o[off + l - 1] = cs[t];
        if (l > 1) {
            const uint32_t poff = (uint32_t)(anc[par] >> 32);
            for (uint32_t j = 0; j + 1 < l; j++) src[off + j] = poff + j;
        }
    }
    __syncthreads();
    for (;;) {                                                        // pointer doubling down to the characters
        bool changed = false;
        for (uint32_t q = tid; q < total32; q += DEC_THREADS) {
            const uint32_t s = src[q];
            if (s != q) {
                const uint32_t s2 = *reinterpret_cast<volatile uint32_t *>(src + s);
                if (s2 != s) { src[q] = s2; changed = true; }
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    for (uint32_t q = tid; q < total32; q += DEC_THREADS) {
        const uint32_t s = src[q];
        if (s != q) o[q] = o[s];                                      // o[s] is a character written above
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t cur = total32;
        if (cur && o[cur - 1] == 0) cur--;                            // mc:3701-3703
        outLen[b] = cur;
    }
}

void lz78_decode_init()
{
    cudaFuncSetAttribute(k78_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(D8Smem));
}

int launch_decode_lz78(const uint8_t *d_file, const uint64_t *d_recOff, uint32_t nBlocks, uint8_t *d_out, size_t outStride,
                       uint32_t *d_outLen, uint8_t *d_scratch, size_t scratchPerBlock, cudaStream_t st)
{
    // grp[] and cs[] start out zero (the reference's calloc'ed arrays, mc:3525 / App. B1)
    cudaMemsetAsync(d_scratch, 0, scratchPerBlock * nBlocks, st);
    k78_decode<<<nBlocks, DEC_THREADS, sizeof(D8Smem), st>>>(d_file, d_recOff, d_out, d_outLen, d_scratch, scratchPerBlock,
                                                                outStride, (uint32_t)std::min<size_t>(outStride, kBlockBytes + 8));
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

} // namespace fcx
