// fcx_huff.cu -- byte-Huffman section encoder (subsystem 3 of the north star) and the
// block / container assembler (subsystem 4).
//
// Replaces my_huffman_encode_char (mc:987-1104) = histogram (mc:997-1000) +
// create_huffman_tree (mc:535-617, stable merge sort mc:458-498) + tree serialisation
// (mc:1021-1066) + huffman_encode_char (mc:849-928), called four times per block by
// my_compress_file_lz77 (mc:2145,2178,2207,2238), and the payload/container layout of
// mc:2137-2243 and mc:4078-4129.
//
// The tree is NOT canonical: bits are the root->leaf path of the reference's own two-queue
// construction (merged node goes after equal weights, mc:588; first popped = left = 0), so
// k_tree replays that construction exactly, one warp per stream.
//
// Bit offsets come from scans, never from a serial bit loop: the histogram pass keeps one
// 256-bin histogram per 8192-symbol chunk, so once code lengths exist the bit length of every
// chunk is a 256-term dot product; a scan over chunks/sections/blocks then gives every chunk
// its absolute bit position in the final container, and k_encode_w writes code words straight
// into the container (the sections sit at odd byte offsets, so everything is addressed as one
// little-endian bit stream over the 4-byte aligned output buffer).
#include "fcx_dev.cuh"
#include "../../include/fcx.h"   // FCX_WITH_HEADER, FCX_FORMAT_X9, header sizes
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace fcx {

// =========================================================================================
// chunk planning
// =========================================================================================
__global__ void __launch_bounds__(1024)
k_plan(Ws ws, uint32_t nBlocks)
{
    __shared__ uint32_t warpSum[32];
    __shared__ uint32_t running;
    const uint32_t nSeg = nBlocks * 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nSeg; base += 1024) {
        const uint32_t s = base + tid;
        uint32_t S = 0, nch = 0, raw = 0;
        if (s < nSeg) {
            const uint32_t b = s >> 2, k = s & 3;
            S = ws.meta[(size_t)b * M_WORDS + M_S0 + k];
            raw = (k == 0 && S <= 1) ? 1u : 0u;              // mc:2095-2107: bitmap of <= 1 byte stays raw
            nch = raw ? 0u : (S + kChunkSyms - 1) / kChunkSyms;
        }
        uint32_t inc = nch;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warpSum[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
        for (int w = 0; w < 32; w++) { uint32_t v = warpSum[w]; if (w < warp) wbase += v; total += v; }
        const uint32_t first = running + wbase + inc - nch;
        if (s < nSeg) {
            uint32_t *si = ws.segInfo + (size_t)s * SI_WORDS;
            si[SI_S] = S; si[SI_FIRST_CHUNK] = first; si[SI_NCHUNKS] = nch; si[SI_RAW] = raw;
            si[SI_R] = 0; si[SI_HDR] = 0; si[SI_W] = 0; si[SI_MAXLEN] = 0;
        }
        __syncthreads();
        if (tid == 0) running += total;
        __syncthreads();
    }
    if (tid == 0) ws.totals[TOT_CHUNKS] = running;
}

__global__ void __launch_bounds__(256)
k_fill_chunks(Ws ws, uint32_t nSeg)
{
    const uint32_t s = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= nSeg) return;
    const uint32_t *si = ws.segInfo + (size_t)s * SI_WORDS;
    const uint32_t first = si[SI_FIRST_CHUNK], nch = si[SI_NCHUNKS];
    for (uint32_t c = lane; c < nch; c += 32) ws.chunkSeg[first + c] = s;
}

// =========================================================================================
// histogram: warp-privatised shared-memory bins, one 256-bin histogram kept per chunk
// =========================================================================================
__global__ void __launch_bounds__(256)
k_hist(Ws ws)
{
    __shared__ uint32_t wh[8][256];
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t nChunks = ws.totals[TOT_CHUNKS];
    for (uint32_t chunk = blockIdx.x; chunk < nChunks; chunk += gridDim.x) {
        const uint32_t s = ws.chunkSeg[chunk];
        const uint32_t *si = ws.segInfo + (size_t)s * SI_WORDS;
        const uint32_t b = s >> 2, k = s & 3;
        const uint32_t c = chunk - si[SI_FIRST_CHUNK];
        const uint32_t S = si[SI_S];
        const uint32_t cnt = min((uint32_t)kChunkSyms, S - c * kChunkSyms);
        const uint8_t *src = ws.strm[k] + (size_t)b * ws.strmStride[k] + (size_t)c * kChunkSyms;
        for (int j = tid; j < 8 * 256; j += 256) (&wh[0][0])[j] = 0;
        __syncthreads();
        uint32_t *h = wh[warp];
        for (uint32_t v = tid; v * 16 < cnt; v += 256) {
            uint4 q = reinterpret_cast<const uint4 *>(src)[v];
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
            const uint32_t valid = min(16u, cnt - v * 16);
            uint32_t cur = w[0] & 0xFF, run = 0;
#pragma unroll
            for (uint32_t j = 0; j < 16; j++) {
                uint32_t sym = (w[j >> 2] >> ((j & 3) * 8)) & 0xFF;
                if (j < valid) {
                    if (sym == cur) run++;
                    else { atomicAdd(&h[cur], run); cur = sym; run = 1; }
                }
            }
            if (run) atomicAdd(&h[cur], run);
        }
        __syncthreads();
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) sum += wh[w][tid];
        ws.chunkHist[(size_t)chunk * 256 + tid] = (uint16_t)sum;
        if (sum) atomicAdd(&ws.segHist[(size_t)s * 256 + tid], sum);
        __syncthreads();
    }
}

// =========================================================================================
// tree + code table + serialised header, one 256-thread CTA per stream (thread = symbol)
// =========================================================================================
struct TreeSmem {
    uint32_t key[256];
    uint32_t sw[256];      // sorted leaf weights
    uint32_t iw[256];      // internal node weights
    uint16_t child[512];   // [2j],[2j+1]: tree index (leaf: symbol, internal: 512-R+j')
    uint16_t parLeaf[256]; // leaf -> 2*parent + side
    uint16_t parInt[256];  // internal node -> 2*parent + side (0xFFFF for the root)
    uint8_t ssym[256];     // sorted leaf symbols
    uint8_t hdr[kSegHdrStride];
    unsigned long long bitsSum;
    uint32_t maxlen, minlen;
};

__global__ void __launch_bounds__(256)
k_tree(Ws ws, uint32_t nSeg, uint32_t escape)
{
    __shared__ TreeSmem sm;
    const int t = threadIdx.x;
    const uint32_t s = blockIdx.x;
    uint32_t *si = ws.segInfo + (size_t)s * SI_WORDS;
    const uint32_t S = si[SI_S];
    if (S == 0 || si[SI_RAW]) return;                       // mc:989-990: empty stream emits nothing (CTA-uniform)

    // ---- stable sort of the non-zero weights: unique key = weight<<8 | symbol, rank by counting ----
    const uint32_t w = ws.segHist[(size_t)s * 256 + t];
    const uint32_t mine = w ? ((w << 8) | (uint32_t)t) : 0xFFFFFFFFu;
    sm.key[t] = mine;
    if (t == 0) { sm.bitsSum = 0; sm.maxlen = 0; sm.minlen = 255; }
    __syncthreads();
    if (w) {
        uint32_t rank = 0;
#pragma unroll 8
        for (int m = 0; m < 256; m++) rank += (sm.key[m] < mine) ? 1u : 0u;
        sm.ssym[rank] = (uint8_t)t;
        sm.sw[rank] = w;
    }
    const uint32_t R = (uint32_t)__syncthreads_count(w != 0);
    const uint32_t Tn = R > 1 ? R - 1 : 0;                   // mc:1013-1019
    const uint32_t mapB = (2 * Tn + 7) / 8;                  // mc:1024-1026
    const uint32_t base = 512 - R;                           // mc:569: first internal node index

    if (t == 0 && R >= 2) {
        // ---- two-queue merge (mc:575-607): on equal weight a leaf, then the older internal node, goes first.
        // The two front weights of either queue live in registers (an exhausted queue reads as INF; real
        // weights are at most 2^20), so the serial chain per pop is a compare and a few moves: the shared
        // loads that refill the look-ahead and the stores that record the tree are off the critical path.
        constexpr uint32_t INF = 0xFFFFFFFFu;
        uint32_t qa = 0, qb = 0, made = 0;
        uint32_t a0 = sm.sw[0], a1 = sm.sw[1];               // R >= 2
        uint32_t b0 = INF, b1 = INF;
        while ((R - qa) + (made - qb) > 1) {
            uint32_t sum = 0;
#pragma unroll
            for (int side = 0; side < 2; side++) {
                if (a0 <= b0) {                              // leaf (both queues cannot be empty here)
                    const uint32_t sym = sm.ssym[qa];
                    sm.child[2 * made + side] = (uint16_t)sym;
                    sm.parLeaf[sym] = (uint16_t)(2 * made + side);
                    sum += a0; qa++;
                    a0 = a1;
                    a1 = qa + 1 < R ? sm.sw[qa + 1] : INF;
                } else {
                    sm.child[2 * made + side] = (uint16_t)(base + qb);
                    sm.parInt[qb] = (uint16_t)(2 * made + side);
                    sum += b0; qb++;
                    b0 = b1;
                    b1 = qb + 1 < made ? sm.iw[qb + 1] : INF;
                }
            }
            sm.iw[made] = sum;
            if (qb == made) b0 = sum;                        // the new node is the front of its queue
            else if (qb + 1 == made) b1 = sum;               // or the one right behind the front
            made++;
        }
        sm.parInt[Tn - 1] = 0xFFFF;                          // root
    }
    __syncthreads();

    // ---- codes: every leaf climbs to the root; bit k of the code = k-th branch from the root, 0 = left
    uint32_t entry = 0, len = 0;
    if (w && R >= 2) {
        // climbing pushes the deepest branch first, so after the last shift the root's branch sits
        // in bit 0 and the deepest in bit len-1: exactly "bit k = k-th branch from the root"
        uint32_t code = 0;
        uint32_t p = sm.parLeaf[t];
        while (p != 0xFFFF && len < 33) {
            code = (code << 1) | (p & 1u);
            len++;
            p = sm.parInt[p >> 1];
        }
        entry = (len >= 1 && len <= 31) ? (code | (1u << len)) : 0u;
    }
    ws.segTab[(size_t)s * 256 + t] = entry;
    if (w) {
        atomicMax(&sm.maxlen, len);
        atomicMin(&sm.minlen, len);
        atomicAdd(&sm.bitsSum, (unsigned long long)w * len);
    }
    // ---- serialised header: Tn | child-is-internal bitmap | (left,right) byte pairs ------------
    if ((uint32_t)t < 1 + mapB) {
        uint32_t v = 0;
        if (t == 0) v = Tn;
        else
            for (int k = 0; k < 8; k++) {
                const uint32_t j = 8 * (t - 1) + k;
                if (j < 2 * Tn && sm.child[j] >= 256) v |= 1u << k;          // mc:1035-1050
            }
        sm.hdr[t] = (uint8_t)v;
    }
    for (uint32_t j = t; j < 2 * Tn; j += 256) {
        const uint32_t ci = sm.child[j];
        sm.hdr[1 + mapB + j] = (uint8_t)(ci >= 256 ? ci - 256 : ci);
    }
    // FCX9 (opt-in format, include/fcx.h): a single-symbol stream names its symbol after the zero node count, so it
    // decodes to what was encoded (the FCX7 section drops it, mc:1013-1019)
    const bool named = escape && R == 1;
    if (t == 0 && named) sm.hdr[1] = sm.ssym[0];
    __syncthreads();
    const uint32_t hdrBytes = 1 + mapB + 2 * Tn + (named ? 1u : 0u);
    for (uint32_t i = t; i < hdrBytes; i += 256) ws.segHdr[(size_t)s * kSegHdrStride + i] = sm.hdr[i];
    if (t == 0) {
        si[SI_R] = R;
        si[SI_HDR] = hdrBytes;
        si[SI_W] = (uint32_t)((sm.bitsSum + 31) >> 5);      // mc:895-923: ceil(totalBits/32) words
        si[SI_MAXLEN] = sm.maxlen | (sm.minlen << 8);           // longest | shortest << 8
        if (sm.maxlen > 31) atomicOr(&ws.totals[TOT_ERR], (uint32_t)E_CODELEN);
    }
}

// =========================================================================================
// per-chunk bit length = dot(chunk histogram, code length)
// =========================================================================================
__global__ void __launch_bounds__(256)
k_chunk_bits(Ws ws)
{
    const int lane = threadIdx.x & 31;
    const uint32_t nChunks = ws.totals[TOT_CHUNKS];
    const uint32_t warpsPerGrid = gridDim.x * 8;
    for (uint32_t chunk = blockIdx.x * 8 + (threadIdx.x >> 5); chunk < nChunks; chunk += warpsPerGrid) {
        const uint32_t s = ws.chunkSeg[chunk];
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int sym = lane + 32 * i;
            const uint32_t e = ws.segTab[(size_t)s * 256 + sym];
            const uint32_t cntv = ws.chunkHist[(size_t)chunk * 256 + sym];
            if (e) bits += cntv * (uint32_t)(31 - __clz(e));
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) bits += __shfl_xor_sync(0xFFFFFFFFu, bits, d);
        if (lane == 0) ws.chunkBits[chunk] = bits;
    }
}

// chunk bit offsets inside each stream (<= 128 chunks per stream: one thread per stream), and the
// 16-byte descriptor k_encode_w reads per chunk (runs after k_block_layout: needs segBitBase)
// chunk descriptor: x = byte offset of the chunk inside its stream arena, z:w = absolute bit
// position of its first code word bit in the container, y = the fields below
constexpr uint32_t CD_CNT_MASK = 0x3FFFu;     // symbols in the chunk (<= 8192)
constexpr int CD_K_SHIFT = 14;                // stream kind 0..3
constexpr int CD_GROUP_SHIFT = 16;            // symbols merged per 32-bit insert: 4 / 3 / 2 / 1; 5 = fixed 8-bit codes (k_encode_w only)
constexpr int CD_SKIP_SHIFT = 19;             // not the fast kernel's: single-symbol stream (no bits) or a slow chunk
constexpr int CD_SEG_SHIFT = 20;              // stream id b*4+k (< 4096)
static_assert(4ull * kMaxBatchBlocks <= (1ull << (32 - CD_SEG_SHIFT)), "stream id of the last block of a batch fits the descriptor");
static_assert((unsigned long long)kMaxBatchBlocks * kSegStride1 < (1ull << 32), "stream arena offsets are 32-bit");

__global__ void __launch_bounds__(256)
k_seg_scan(Ws ws, uint32_t nSeg)
{
    // one warp per stream: 32 chunks per trip, warp prefix sum of their bit lengths
    const int lane = threadIdx.x & 31;
    const uint32_t s = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= nSeg) return;
    const uint32_t *si = ws.segInfo + (size_t)s * SI_WORDS;
    const uint32_t first = si[SI_FIRST_CHUNK], nch = si[SI_NCHUNKS], S = si[SI_S], W = si[SI_W];
    const uint32_t maxlen = si[SI_MAXLEN] & 0xFFu, minlen = si[SI_MAXLEN] >> 8;
    const bool longCodes = maxlen > 24;                      // fast path keeps a code in 24 bits (GROUP 1) or merges 2..4 into 32
    // 5: every code is exactly 8 bits (flat histogram over all 256 bytes: incompressible input), a byte translation
    const uint32_t group = (maxlen == 8 && minlen == 8) ? 5u : maxlen <= 8 ? 4u : maxlen <= 10 ? 3u : maxlen <= 16 ? 2u : 1u;
    const uint32_t b = s >> 2, k = s & 3;
    const unsigned long long bitBase = ws.segBitBase[s];
    const uint32_t arena = b * ws.strmStride[k];
    uint32_t running = 0, slow = 0;
    for (uint32_t c0 = 0; c0 < nch; c0 += 32) {
        const uint32_t c = c0 + lane;
        const bool have = c < nch;
        const uint32_t bits = have ? ws.chunkBits[first + c] : 0u;
        uint32_t inc = bits;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        const uint32_t off = running + inc - bits;
        if (have) {
            ws.chunkBitOff[first + c] = off;
            const bool sl = W != 0 && longCodes;                          // k_encode_slow's
            ws.chunkSlow[first + c] = sl ? 1 : 0;
            slow += sl ? 1u : 0u;
            const uint32_t cnt = min((uint32_t)kChunkSyms, S - c * kChunkSyms);
            const unsigned long long gb0 = bitBase + off;
            const uint32_t skip = (sl || W == 0) ? 1u : 0u;
            ws.chunkDesc[first + c] = make_uint4(arena + c * kChunkSyms,
                                                 cnt | (k << CD_K_SHIFT) | (group << CD_GROUP_SHIFT) | (skip << CD_SKIP_SHIFT) | (s << CD_SEG_SHIFT),
                                                 (uint32_t)gb0, (uint32_t)(gb0 >> 32));
        }
        running += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) slow += __shfl_xor_sync(0xFFFFFFFFu, slow, d);
    if (lane == 0) {
        if (slow) atomicAdd(&ws.totals[TOT_SLOW], slow);
        if (nch && ((running + 31) >> 5) != W) atomicOr(&ws.totals[TOT_ERR], (uint32_t)E_INTERNAL);
        // what the stream costs k_encode_w, in units of 64 symbols of the cheapest variable-length path: the
        // pack pass grows with the number of groups per run (8 / 11 / 16 / 32), the fixed path has none
        const uint32_t weight = group == 5 ? 7u : group == 4 ? 16u : group == 3 ? 18u : group == 2 ? 21u : 30u;
        ws.segCost[s] = (W == 0 || longCodes || nch == 0) ? 0u : (S / 64 + 2 * nch) * weight;
    }
}

// Contiguous chunk ranges of equal estimated cost, one per k_encode_w CTA: a range of ten chunks of a 16-bit
// stream takes a third longer than ten chunks of an 8-bit one, and on small inputs a CTA sees only one or two
// streams.  One CTA: exclusive prefix sum of the per-stream costs, then every boundary j finds the stream
// holding cost j/G of the total and cuts it proportionally (chunks of one stream cost the same).
__global__ void __launch_bounds__(1024)
k_enc_split(Ws ws, uint32_t nSeg, uint32_t G)
{
    constexpr uint32_t kPreSmem = 4096;                        // prefix kept in shared memory up to 1024 blocks per batch
    __shared__ uint32_t preS[kPreSmem + 1];
    __shared__ uint32_t warpSum[32];
    __shared__ uint32_t running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool inSmem = nSeg <= kPreSmem;
    // a stream costs at most ~5e5 units: 4096 of them fit 32 bits as they are, larger batches are scaled down
    const uint32_t scale = inSmem ? 0u : 7u;
    if (tid == 0) running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nSeg; base += 1024) {
        const uint32_t s = base + tid;
        const uint32_t v = s < nSeg ? (ws.segCost[s] >> scale) : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warpSum[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
        for (int w = 0; w < 32; w++) { const uint32_t x = warpSum[w]; if (w < warp) wbase += x; total += x; }
        if (s < nSeg) {
            ws.segCost[s] = running + wbase + inc - v;
            if (inSmem) preS[s] = running + wbase + inc - v;
        }
        __syncthreads();
        if (tid == 0) running += total;
        __syncthreads();
    }
    const uint32_t total = running;
    if (tid == 0) { ws.segCost[nSeg] = total; if (inSmem) preS[nSeg] = total; }
    const uint32_t *pre = inSmem ? preS : ws.segCost;
    __threadfence_block();
    __syncthreads();
    const uint32_t nChunks = ws.totals[TOT_CHUNKS];
    for (uint32_t j = tid; j <= G; j += 1024) {
        uint32_t cut = nChunks;
        if (j == 0) cut = 0;
        else if (j < G && total) {
            const uint32_t target = (uint32_t)((unsigned long long)total * j / G);
            uint32_t lo = 0, hi = nSeg;                        // last stream whose prefix is <= target
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (pre[mid] <= target) lo = mid; else hi = mid; }
            const uint32_t before = pre[lo], cost = pre[lo + 1] - before;
            const uint32_t *si = ws.segInfo + (size_t)lo * SI_WORDS;
            const uint32_t first = si[SI_FIRST_CHUNK], nch = si[SI_NCHUNKS];
            cut = first + (cost ? min(nch, (uint32_t)((unsigned long long)nch * (target - before) / cost)) : 0u);
        } else if (j < G)
            cut = (uint32_t)((unsigned long long)nChunks * j / G);     // nothing to encode: any split
        ws.encSplit[j] = cut;
    }
}

// =========================================================================================
// payload sizes -> record offsets -> absolute bit position of every stream's code words
// =========================================================================================
__device__ __forceinline__ uint32_t section_bytes(const uint32_t *si)
{
    const uint32_t S = si[SI_S];
    if (S == 0) return 0;
    if (si[SI_RAW]) return S;
    return si[SI_HDR] + 4 + 4 * si[SI_W];
}
static_assert(SI_WORDS == 8, "a segInfo row is read as two 16-byte vectors");
// the same from a row held in registers (lo = words 0..3, hi = words 4..7)
__device__ __forceinline__ uint32_t section_bytes(const uint4 &lo, const uint4 &hi)
{
    if (lo.x == 0) return 0;                 // SI_S
    if (hi.z) return lo.x;                   // SI_RAW
    return lo.z + 4 + 4 * lo.w;              // SI_HDR + 4 + 4 * SI_W
}

__global__ void __launch_bounds__(1024)
k_block_layout(Ws ws, uint32_t nBlocks, uint32_t headerLen, uint64_t outCap)
{
    __shared__ unsigned long long warpSum[32];
    __shared__ unsigned long long running;
    __shared__ unsigned long long accSin, accSout, accTok, accMat;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        running = ws.counters[C_CURSOR] + headerLen;
        accSin = accSout = accTok = accMat = 0;
    }
    __syncthreads();
    for (uint32_t base = 0; base < nBlocks; base += 1024) {
        const uint32_t b = base + tid;
        unsigned long long rec = 0;
        uint32_t sec[4] = {0, 0, 0, 0}, hdrB[4] = {0, 0, 0, 0};
        if (b < nBlocks) {
            // the block's four segInfo rows: eight independent 16-byte loads, one round trip
            uint4 lo[4], hi[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint4 *row = reinterpret_cast<const uint4 *>(ws.segInfo + ((size_t)b * 4 + k) * SI_WORDS);
                lo[k] = row[0]; hi[k] = row[1];
            }
            const uint32_t T = ws.meta[(size_t)b * M_WORDS + M_T], Pc = ws.meta[(size_t)b * M_WORDS + M_PC];
            unsigned long long sin = 0, sout = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                sec[k] = section_bytes(lo[k], hi[k]);
                hdrB[k] = lo[k].z;
                if (lo[k].w) { sin += lo[k].x; sout += 4ull * lo[k].w; }
            }
            rec = 4ull + 4 + sec[0] + sec[1] + 4 + sec[2] + 4 + sec[3];   // u32 size | payload (mc:4111-4114)
            atomicAdd(&accSin, sin); atomicAdd(&accSout, sout);
            atomicAdd(&accTok, (unsigned long long)T);
            atomicAdd(&accMat, (unsigned long long)Pc);
        }
        unsigned long long inc = rec;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warpSum[warp] = inc;
        __syncthreads();
        unsigned long long wbase = 0, total = 0;
        for (int w = 0; w < 32; w++) { unsigned long long v = warpSum[w]; if (w < warp) wbase += v; total += v; }
        const unsigned long long off = running + wbase + inc - rec;
        if (b < nBlocks) {
            ws.blkOff[b] = off;
            // absolute bit position of word 0 of each stream's code words
            unsigned long long p = off + 4 + 4;                                  // size, T
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k == 2 || k == 3) p += 4;                                    // Pc before p section, G before Golomb section
                ws.segBitBase[(size_t)b * 4 + k] = 8ull * (p + hdrB[k] + 4);
                p += sec[k];
            }
        }
        __syncthreads();
        if (tid == 0) running += total;
        __syncthreads();
    }
    if (tid == 0) {
        ws.blkOff[nBlocks] = running;
        ws.counters[C_LAST_BYTES] = running - ws.counters[C_CURSOR];
        ws.counters[C_CURSOR] = running;
        ws.counters[C_SIN] += accSin; ws.counters[C_SOUT] += accSout;
        ws.counters[C_TOKENS] += accTok; ws.counters[C_MATCHES] += accMat;
        ws.counters[C_BLOCKS] += nBlocks;
        if (running > outCap) atomicOr(&ws.totals[TOT_ERR], (uint32_t)E_CAPACITY);
    }
}

// =========================================================================================
// encode: the HBM-bound kernel (algorithmic bytes = S_in + S_out)
// =========================================================================================
// k_encode_w (warp per sub-chunk, no barriers in the chunk loop) is the encoder; k_encode_slow takes the
// rare streams whose codes are longer than 24 bits.  Common ground: a thread owns a run of 32 consecutive
// symbols (two 16-byte loads), looks every symbol up once in a shared-memory table replicated per lane
// (conflict-free), a scan of the run lengths gives every run its bit position, and the codes are packed
// from registers through a 64-bit accumulator whose low word is flushed to a shared staging buffer.  A run
// is >= 32 bits, so a staging word has at most two writers: the tail of run t travels by shuffle into the
// first word of run t+1 -- no atomics inside the stage.  The staged words go out coalesced; only the
// first/last word of a (sub-)chunk (shared with its neighbour or with header bytes) is merged with atomicOr
// on the pre-zeroed container.
constexpr int ENC_THREADS = 256;
constexpr int ENC_RUN = kChunkSyms / ENC_THREADS;     // 32 symbols per thread
static_assert(ENC_RUN == 32, "run length is baked into the carry logic (>= 32 bits per run)");
constexpr int ENC_STAGE_WORDS = 3064;                 // k_encode_slow's CTA-wide stage; denser chunks go straight to the container
constexpr int ENC_DESC_BATCH = 64;

struct EncSmem {
    uint32_t tab[256 * 32];            // code | 1<<len, one copy per lane
    alignas(16) uint32_t stage[ENC_STAGE_WORDS + 8];
    uint32_t warpBits[8];
    uint32_t warpTail[8];
};

__device__ __forceinline__ uint32_t enc_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t enc_lds(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void enc_sts(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 enc_lds64(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

// generic path, sentinel entries (code | 1<<len, len <= 31): one 64-bit insert per symbol
__device__ __forceinline__ void enc_pack_generic(const uint32_t (&ent)[32], uint32_t p0, uint32_t stageA,
                                                 uint32_t &tail, uint32_t &tailIdx)
{
    unsigned long long acc = 0;
    uint32_t accBits = p0 & 31, w = stageA + 4 * (p0 >> 5);
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const uint32_t e = ent[j];
        const uint32_t len = (uint32_t)(31 - __clz(e));
        acc |= (unsigned long long)(e ^ (1u << len)) << accBits;
        accBits += len;
        if (accBits >= 32) { enc_sts(w, (uint32_t)acc); w += 4; acc >>= 32; accBits -= 32; }
    }
    tail = (uint32_t)acc; tailIdx = (w - stageA) >> 2;
}

// Optional stage swizzle at 16-byte granularity (unit u lives at u ^ ((u >> 3) & 7)): runs of equal length
// would otherwise put every lane of a store on the same four banks.  Whole units stay intact, so the
// copy-out still moves 128 bits at a time.  (See FCX_ENCW_SWIZZLE below.)
__device__ __forceinline__ uint32_t enc_swz_byte(uint32_t byteOff) { return byteOff ^ ((byteOff >> 3) & 0x70u); }
__device__ __forceinline__ uint32_t enc_swz_word(uint32_t k) { return k ^ (((k >> 5) & 7u) << 2); }
__device__ __forceinline__ uint32_t enc_swz_unit(uint32_t v) { return v ^ ((v >> 3) & 7u); }

// ---- warp-autonomous fast kernel ------------------------------------------------------------
// The unit of work is a WARP: warp w of the
// CTA owns sub-chunk w (1024 symbols, 32 per lane) of every chunk in the CTA's range, with a
// private 2 KB stage.  There is no CTA barrier in the chunk loop (only when the stream -- the code
// table -- changes and once per 64 descriptors): the bit position of a sub-chunk is the chunk's
// position plus the totals of the warps before it, handed from warp to warp through one tagged
// shared word per (chunk, warp) (decoupled look-back inside the CTA; warp w never runs ahead of
// warp w-1, so the spin cannot deadlock).  Symbols come straight from global memory into
// registers, one chunk ahead.  Sub-chunk boundaries fall inside container words, so the first and
// last word of a sub-chunk are merged with atomicOr on the pre-zeroed container (two per 1024
// symbols); everything between goes out as 128-bit stores.  A sub-chunk too dense for the stage
// (> 16 bits per symbol on average; adversarial) is written with atomicOr straight from registers.
constexpr int EW_WARPS = 8;
#ifndef FCX_ENCW_CTAS
#define FCX_ENCW_CTAS 3                  // resident CTAs per SM: 80 registers per thread, nothing spilled (4 CTAs: 64 and spills in the chunk loop; measured slower)
#endif
constexpr int EW_CTAS = FCX_ENCW_CTAS;
constexpr int EW_THREADS = EW_WARPS * 32;
static_assert(EW_THREADS == 256, "one thread per symbol when the code table is rebuilt");
static_assert(kChunkSyms == EW_WARPS * 32 * ENC_RUN, "a chunk is one sub-chunk per warp");
#ifndef FCX_ENCW_STAGE_WORDS
#define FCX_ENCW_STAGE_WORDS 544      // tests build a variant with a tiny stage to exercise the direct path
#endif
constexpr int EW_STAGE_WORDS = FCX_ENCW_STAGE_WORDS;
static_assert(EW_STAGE_WORDS % 32 == 0, "the stage swizzle permutes inside aligned 32-word groups");
constexpr uint32_t EW_STAGE_BITS = (uint32_t)EW_STAGE_WORDS * 32u - 32u;
constexpr uint32_t EW_PREF_SHIFT = 18;                 // low 18 bits: bit count of one sub-chunk
static_assert((uint32_t)kChunkSyms * 24u < (1u << EW_PREF_SHIFT), "a chunk's bit count fits the prefix field");

// Stage swizzle of k_encode_w.  Off by default: the streams whose runs all
// have the same length -- the case the swizzle exists for -- are the fixed 8-bit ones, and those no longer
// touch the stage; without it every flush is three instructions shorter.
#ifndef FCX_ENCW_SWIZZLE
#define FCX_ENCW_SWIZZLE 0
#endif
__device__ __forceinline__ uint32_t ew_swz_byte(uint32_t byteOff) { return FCX_ENCW_SWIZZLE ? enc_swz_byte(byteOff) : byteOff; }
__device__ __forceinline__ uint32_t ew_swz_word(uint32_t k) { return FCX_ENCW_SWIZZLE ? enc_swz_word(k) : k; }
__device__ __forceinline__ uint32_t ew_swz_unit(uint32_t v) { return FCX_ENCW_SWIZZLE ? enc_swz_unit(v) : v; }

struct EncWSmem {
    alignas(1024) uint32_t stage[EW_WARPS][EW_STAGE_WORDS];
    alignas(16) unsigned long long tab[256 * 16];      // tab[sym*16 + (lane & 15)] = code | (u64)(1 << len) << 32
    alignas(16) uint4 desc[ENC_DESC_BATCH];            // the batch's descriptors without the skipped chunks
    uint32_t pref[ENC_DESC_BATCH][EW_WARPS];           // tag << 18 | bit count of the sub-chunk, one word per (chunk of the batch, warp)
    uint32_t tabNext[256];                             // the next stream's code table entries (code | 1<<len), fetched a chunk ahead
    uint32_t nDesc, curSeg, tag;                       // loop state every warp re-reads per chunk (registers are scarce: 64 per thread)
};

__device__ __forceinline__ uint4 encw_ldg128(const uint8_t *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t encw_ldg32(const uint8_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

constexpr uint32_t EW_FIXED8 = 5;     // descriptor group code: every code of the stream is exactly 8 bits

// This thread's symbols of the chunk at src (y = the descriptor's flag word).  Variable-length streams:
// the 32 consecutive symbols of run `tid` (nothing when the chunk has fewer runs).  Fixed 8-bit streams:
// words lane, lane + 32, ... of the warp's 1024-symbol sub-chunk, so loads and stores are both coalesced.
__device__ __forceinline__ void encw_fetch(uint32_t (&sy)[8], const uint8_t *src, uint32_t y, uint32_t tid)
{
    const uint32_t cnt = y & CD_CNT_MASK;
    if (((y >> CD_GROUP_SHIFT) & 7u) == EW_FIXED8) {
        const uint32_t b0 = 4u * ((tid & ~31u) * 8u + (tid & 31u));      // byte offset of word (256 * warp + lane)
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (b0 + 128u * j < cnt) sy[j] = encw_ldg32(src + b0 + 128u * j);   // may read <= 3 bytes past cnt (arena slack)
    } else if (tid < cnt / ENC_RUN) {
        const uint4 a = encw_ldg128(src + tid * 32), b = encw_ldg128(src + tid * 32 + 16);
        sy[0] = a.x; sy[1] = a.y; sy[2] = a.z; sy[3] = a.w; sy[4] = b.x; sy[5] = b.y; sy[6] = b.z; sy[7] = b.w;
    }
}

// Fixed 8-bit codes: the code words of the stream are a byte translation of its symbols, so nothing is
// scanned, staged or handed over.  Lane l owns words l, l + 32, ... of the sub-chunk; a word of four codes is
// shifted to the container's byte alignment with its predecessor (one shuffle) and stored coalesced.  Only the
// first and last word of the sub-chunk can be shared with a neighbour (atomicOr).  Covers the ragged end too.
__device__ __forceinline__ void encw_fixed8(uint32_t *__restrict__ out32, uint32_t (&sy)[8], uint32_t cnt,
                                            const uint8_t *nextSrc, uint32_t nextY, uint32_t tid,
                                            unsigned long long gb0, uint32_t tab32Lane, int warp, int lane)
{
    const uint32_t n = cnt > (uint32_t)warp * 1024u ? min(1024u, cnt - (uint32_t)warp * 1024u) : 0u;   // symbols of this sub-chunk
    // this path is all loads and stores: the next chunk's symbols are requested before anything else
    uint32_t c[8];
#pragma unroll
    for (int j = 0; j < 8; j++) c[j] = sy[j];
    if (nextSrc) encw_fetch(sy, nextSrc, nextY, tid);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t m = 4u * (32u * j + (uint32_t)lane);               // first symbol of the word
        uint32_t v = 0;
        if (m < n) {
            const uint32_t w = c[j];
            v = enc_lds(tab32Lane + (__byte_perm(w, 0u, 0x4440u) << 7));
            v += enc_lds(tab32Lane + (__byte_perm(w, 0u, 0x4441u) << 7)) << 8;
            v += enc_lds(tab32Lane + (__byte_perm(w, 0u, 0x4442u) << 7)) << 16;
            v += enc_lds(tab32Lane + (__byte_perm(w, 0u, 0x4443u) << 7)) << 24;
            if (m + 4u > n) v &= (1u << (8u * (n - m))) - 1u;             // bytes past the end of the stream
        }
        c[j] = v;
    }
    if (n == 0) return;                                                   // warp-uniform
    const unsigned long long B = (gb0 >> 3) + (unsigned long long)warp * 1024u;    // container byte of the first code
    const uint32_t a = (uint32_t)B & 3u, sh = 8u * a;
    uint32_t *dst = out32 + (B >> 2);
    uint32_t carry = 0;                                                   // the word before this round's lane 0
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, c[j], 1);
        const uint32_t last = __shfl_sync(0xFFFFFFFFu, c[j], 31);
        if (lane == 0) prev = carry;
        carry = last;
        const uint32_t v = __funnelshift_l(prev, c[j], sh);              // (c << sh) | (prev >> (32 - sh))
        const uint32_t k = 32u * j + (uint32_t)lane;
        if ((k > 0 || a == 0) && 4u * k + 4u <= a + n) dst[k] = v;        // all four bytes are this sub-chunk's
        else if (v) atomicOr(&dst[k], v);
    }
    if (a && lane == 0) {
        const uint32_t v = carry >> (32u - sh);                           // word 256: the spill of a full sub-chunk
        if (v) atomicOr(&dst[256], v);
    }
}

// One sub-chunk.  sy holds this lane's symbols on entry; once they are looked up it is refilled with the
// lane's symbols of the next chunk (nextSrc / nextY), which then travel while this one is packed and
// copied out.  Returns the chunk's bit count up to and including this warp.
template <int GROUP>
__device__ __forceinline__ uint32_t encw_sub(uint32_t *__restrict__ out32, uint32_t (&sy)[8], uint32_t run, uint32_t nRuns,
                                             const uint8_t *nextSrc, uint32_t nextY,
                                             unsigned long long gb0, uint32_t tabLane, uint32_t stageA,
                                             uint32_t prefA, uint32_t tag, int warp, int lane)
{
    constexpr int NG = (32 + GROUP - 1) / GROUP;
    const bool active = run < nRuns;
    uint32_t g[NG], gl[GROUP == 1 ? 1 : NG];
    uint32_t myBits = 0;
    if (active) {
#pragma unroll
        for (int q = 0; q < NG; q++) {
            // entries are {code, 2^len}: a code is shifted into place by a multiply-add with the running
            // power of two (FMA pipe; the ALU pipe is the busy one), and the group's length falls out of
            // popc(2^len - 1) -- which is also right when four 8-bit codes wrap the power to 0
            uint32_t acc = 0, pw = 1;
#pragma unroll
            for (int k = 0; k < GROUP; k++) {
                const int j = q * GROUP + k;
                if (j < 32) {
                    const uint32_t sym = __byte_perm(sy[j >> 2], 0u, 0x4440u + (j & 3));
                    const uint2 e = enc_lds64(tabLane + (sym << 7));
                    acc = k == 0 ? e.x : e.x * pw + acc;
                    pw = k == 0 ? e.y : pw * e.y;
                }
            }
            const uint32_t len = (uint32_t)__popc(pw - 1u);
            if (GROUP == 1) g[q] = acc | (len << 24);
            else { g[q] = acc; gl[q] = len; }
            myBits += len;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NG; q++) { g[q] = 0; if (GROUP != 1) gl[q] = 0; }
    }
    if (nextSrc) encw_fetch(sy, nextSrc, nextY, run);
    // ---- warp scan, then the hand-over of the running bit count from the warp before ----------
    uint32_t inc = myBits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    // every warp publishes its own total as soon as it has it (no chain); lanes 0..warp-1 of this warp wait for
    // the words of the warps before it and the sum of those totals is where this sub-chunk starts
    if (lane == 0)
        asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(prefA + 4u * (uint32_t)warp), "r"(total | (tag << EW_PREF_SHIFT)) : "memory");
    uint32_t before = 0;
    if (warp > 0) {
        uint32_t v;
        bool ok;
        do {
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(prefA + 4u * (uint32_t)(lane & (EW_WARPS - 1))) : "memory");
            ok = lane >= warp || (v >> EW_PREF_SHIFT) == tag;
        } while (!__all_sync(0xFFFFFFFFu, ok));
        before = __reduce_add_sync(0xFFFFFFFFu, lane < warp ? (v & ((1u << EW_PREF_SHIFT) - 1u)) : 0u);
    }
    const uint32_t through = before + total;
    if (nRuns <= (uint32_t)warp * 32u) return through;   // warp-uniform: nothing of this chunk is ours
    const uint32_t nAct = min(32u, nRuns - (uint32_t)warp * 32u);

    // stage word k holds container word wBase + k (wBase a multiple of 4: 16-byte units line up)
    const unsigned long long gbW = gb0 + before;
    const uint32_t r0 = (uint32_t)gbW & 127u;
    uint32_t *dst = out32 + ((gbW >> 5) & ~3ull);
    const uint32_t p0 = r0 + inc - myBits;

    if (r0 + total > EW_STAGE_BITS) {
        // ---- too dense for the stage: straight to the container -----------------------------------
        if (active) {
            uint32_t bp = p0;
#pragma unroll
            for (int q = 0; q < NG; q++) {
                const uint32_t bitsq = GROUP == 1 ? (g[q] & 0xFFFFFFu) : g[q];
                const uint32_t lenq = GROUP == 1 ? (g[q] >> 24) : gl[q];
                const uint32_t lo = bitsq << (bp & 31), hi = __funnelshift_l(bitsq, 0u, bp);
                if (lo) atomicOr(&dst[bp >> 5], lo);
                if (hi) atomicOr(&dst[(bp >> 5) + 1], hi);
                bp += lenq;
            }
        }
        return through;
    }

    // ---- pack: insert the groups, flush the low word of the 64-bit accumulator -------------------
    uint32_t tail = 0, tailIdx = 0;
    if (active) {
        uint32_t lo = 0, accBits = p0 & 31, w = 4 * (p0 >> 5);
#pragma unroll
        for (int q = 0; q < NG; q++) {
            const uint32_t bitsq = GROUP == 1 ? (g[q] & 0xFFFFFFu) : g[q];
            const uint32_t lenq = GROUP == 1 ? (g[q] >> 24) : gl[q];
            lo |= bitsq << accBits;
            const uint32_t hi = __funnelshift_l(bitsq, 0u, accBits);     // bits that spill past bit 31 (0 when accBits == 0)
            accBits += lenq;
            if (accBits >= 32) { enc_sts(stageA + ew_swz_byte(w), lo); w += 4; lo = hi; }
            accBits &= 31;
        }
        tail = lo; tailIdx = w >> 2;
    }
    // the tail of run t-1 lives in the low bits of the head word of run t
    const uint32_t prevTail = __shfl_up_sync(0xFFFFFFFFu, tail, 1);
    if (active) {
        const uint32_t headA = stageA + 4 * ew_swz_word(p0 >> 5);
        if (lane > 0) enc_sts(headA, enc_lds(headA) | prevTail);                          // own word, own earlier store
        if ((uint32_t)lane == nAct - 1) enc_sts(stageA + 4 * ew_swz_word(tailIdx), tail);   // nobody follows: park the tail
    }
    __syncwarp();
    // ---- copy out: whole 16-byte units as 128-bit stores, the two shared edge words by atomicOr ----
    const uint32_t first = r0 >> 5, fbit = r0 & 31;
    const uint32_t endBit = r0 + total, last = endBit >> 5, rem = endBit & 31;
    const uint32_t vBeg = (first + (fbit ? 1u : 0u) + 3u) >> 2, vEnd = last >> 2;
    if (vBeg < vEnd) {
        for (uint32_t v = vBeg + lane; v < vEnd; v += 32) {
            uint4 q;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(stageA + 16 * ew_swz_unit(v)));
            reinterpret_cast<uint4 *>(dst)[v] = q;
        }
        if (lane < 8) {                                              // scalar edges
            const uint32_t k = lane < 4 ? first + lane : 4 * vEnd + (lane - 4);
            const bool inHead = lane < 4 && k < 4 * vBeg, inTail = lane >= 4 && k < last;
            if (inHead || inTail) {
                const uint32_t v = enc_lds(stageA + 4 * ew_swz_word(k));
                if (k == first && fbit) atomicOr(&dst[k], v); else dst[k] = v;
            }
        }
    } else {
        for (uint32_t k = first + lane; k < last; k += 32) {
            const uint32_t v = enc_lds(stageA + 4 * ew_swz_word(k));
            if (k == first && fbit) atomicOr(&dst[k], v); else dst[k] = v;
        }
    }
    if (lane == 0 && rem) atomicOr(&dst[last], enc_lds(stageA + 4 * ew_swz_word(last)));
    __syncwarp();                                       // the stage is rewritten by the next sub-chunk
    return through;
}

__global__ void __launch_bounds__(EW_THREADS, EW_CTAS)
k_encode_w(Ws ws, uint32_t *__restrict__ out32)
{
    extern __shared__ __align__(1024) unsigned char enc_smem_raw[];
    EncWSmem &sm = *reinterpret_cast<EncWSmem *>(enc_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint2 tot = *reinterpret_cast<const uint2 *>(ws.totals);      // TOT_CHUNKS, TOT_ERR
    static_assert(TOT_CHUNKS == 0 && TOT_ERR == 1, "read as one 8-byte vector");
    if (tot.y) return;                                  // capacity / internal error: write nothing
    const uint32_t nChunks = tot.x;
    const uint32_t cBeg = ws.encSplit[blockIdx.x], cEnd = min(nChunks, ws.encSplit[blockIdx.x + 1]);   // k_enc_split
    const uint32_t tabLane = enc_smem_addr(sm.tab) + 8 * (lane & 15);
    const uint32_t stageA = enc_smem_addr(sm.stage[warp]);
    for (int i = tid; i < ENC_DESC_BATCH * EW_WARPS; i += EW_THREADS) (&sm.pref[0][0])[i] = 0;   // tag 0 = never published
    uint32_t batchNo = 0;
    if (tid == 0) sm.curSeg = 0xFFFFFFFFu;
    uint32_t sy[8];
#pragma unroll
    for (int q = 0; q < 8; q++) sy[q] = 0;
    for (uint32_t base = cBeg; base < cEnd; base += ENC_DESC_BATCH, batchNo++) {
        __syncthreads();                                // everyone is done with the previous batch of descriptors / hand-over words
        if (tid == 0) sm.tag = batchNo % 16383u + 1u;   // never 0; a hand-over word is written once per batch
        // ---- the batch's descriptors, skipped chunks (single-symbol streams, k_encode_slow's) squeezed out
        if (warp < 2) {
            const uint32_t c0 = base + lane, c1 = base + 32 + lane;
            const uint4 mine = (warp ? c1 : c0) < cEnd ? ws.chunkDesc[warp ? c1 : c0] : make_uint4(0, 1u << CD_SKIP_SHIFT, 0, 0);
            const uint32_t y0 = warp == 0 ? mine.y : (c0 < cEnd ? ws.chunkDesc[c0].y : (1u << CD_SKIP_SHIFT));
            const uint32_t keep0 = __ballot_sync(0xFFFFFFFFu, !((y0 >> CD_SKIP_SHIFT) & 1u));
            const uint32_t keepMine = warp == 0 ? keep0 : __ballot_sync(0xFFFFFFFFu, !((mine.y >> CD_SKIP_SHIFT) & 1u));
            const uint32_t pos = (warp ? __popc(keep0) : 0u) + __popc(keepMine & ((1u << lane) - 1u));
            if ((keepMine >> lane) & 1u) sm.desc[pos] = mine;
            if (warp == 1 && lane == 0) sm.nDesc = __popc(keep0) + __popc(keepMine);
        }
        __syncthreads();
        const uint32_t nDescA = enc_smem_addr(&sm.nDesc), curSegA = enc_smem_addr(&sm.curSeg), tagA = enc_smem_addr(&sm.tag);
        if (enc_lds(nDescA)) {                          // the first chunk of a batch is fetched in the open
            const uint4 d = sm.desc[0];
            encw_fetch(sy, ws.strm[(d.y >> CD_K_SHIFT) & 3u] + d.x, d.y, (uint32_t)tid);
        }
        for (uint32_t i = 0; i < enc_lds(nDescA); i++) {
            const uint4 d = sm.desc[i];
            const uint32_t cnt = d.y & CD_CNT_MASK, group = (d.y >> CD_GROUP_SHIFT) & 7u, s = d.y >> CD_SEG_SHIFT;
            const unsigned long long gb0 = (unsigned long long)d.z | ((unsigned long long)d.w << 32);
            const uint32_t nRuns = cnt / ENC_RUN;
            if (s != enc_lds(curSegA)) {
                // EW_THREADS == 256 symbols; code | 1<<len.  Fetched one chunk ahead (below) unless this is the first chunk of a batch.
                uint32_t e;
                if (i) {
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    e = sm.tabNext[tid];
                } else
                    e = ws.segTab[(size_t)s * 256 + tid];
                const uint32_t len = e ? (uint32_t)(31 - __clz(e)) : 0u;
                const uint32_t code = e ? (e ^ (1u << len)) : 0u;
                __syncthreads();                        // every warp is done with the previous stream's table
                if (group == EW_FIXED8) {               // 32-bit entries (the code byte), one copy per lane, same 128-byte rows
                    const uint4 q = make_uint4(code, code, code, code);
#pragma unroll
                    for (int c = 0; c < 8; c++) reinterpret_cast<uint4 *>(sm.tab)[tid * 8 + ((c + lane) & 7)] = q;
                } else {
                    const unsigned long long ent = (unsigned long long)code | ((unsigned long long)(1u << len) << 32);
#pragma unroll 4
                    for (int c = 0; c < 16; c++) sm.tab[tid * 16 + ((c + lane) & 15)] = ent;
                }
                if (tid == 0) sm.curSeg = s;
                __syncthreads();
            }
            const uint8_t *nextSrc = nullptr;
            uint32_t nextY = 0;
            if (i + 1 < enc_lds(nDescA)) {
                const uint2 dn = *reinterpret_cast<const uint2 *>(&sm.desc[i + 1]);
                nextSrc = ws.strm[(dn.y >> CD_K_SHIFT) & 3u] + dn.x;
                nextY = dn.y;
                const uint32_t sn = dn.y >> CD_SEG_SHIFT;
                if (sn != s)                            // the next stream's table entry of this thread, parked in shared memory
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(enc_smem_addr(&sm.tabNext[tid])), "l"(ws.segTab + (size_t)sn * 256 + tid) : "memory");
            }
            if (group == EW_FIXED8) {
                encw_fixed8(out32, sy, cnt, nextSrc, nextY, (uint32_t)tid, gb0, enc_smem_addr(sm.tab) + 4 * lane, warp, lane);
                continue;
            }
            const uint32_t prefA = enc_smem_addr(&sm.pref[i][0]), tag = enc_lds(tagA);
            const uint32_t through =
                group == 4   ? encw_sub<4>(out32, sy, (uint32_t)tid, nRuns, nextSrc, nextY, gb0, tabLane, stageA, prefA, tag, warp, lane)
                : group == 3 ? encw_sub<3>(out32, sy, (uint32_t)tid, nRuns, nextSrc, nextY, gb0, tabLane, stageA, prefA, tag, warp, lane)
                : group == 2 ? encw_sub<2>(out32, sy, (uint32_t)tid, nRuns, nextSrc, nextY, gb0, tabLane, stageA, prefA, tag, warp, lane)
                             : encw_sub<1>(out32, sy, (uint32_t)tid, nRuns, nextSrc, nextY, gb0, tabLane, stageA, prefA, tag, warp, lane);
            // ---- ragged end of the stream (< 32 symbols): one symbol per lane of the warp that knows the chunk's bit count
            if (warp == EW_WARPS - 1 && cnt > nRuns * ENC_RUN) {
                const uint32_t t = nRuns * ENC_RUN + (uint32_t)lane;
                uint32_t code = 0, len = 0;
                if (t < cnt) {
                    const uint8_t *src = ws.strm[(d.y >> CD_K_SHIFT) & 3u] + d.x;
                    const uint2 e = enc_lds64(tabLane + ((uint32_t)src[t] << 7));
                    code = e.x; len = (uint32_t)__popc(e.y - 1u);
                }
                uint32_t inc = len;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, dd);
                    if (lane >= dd) inc += u;
                }
                const unsigned long long bp = gb0 + through + (inc - len);
                const uint32_t lo = code << (bp & 31), hi = __funnelshift_l(code, 0u, (uint32_t)bp);
                if (lo) atomicOr(&out32[bp >> 5], lo);
                if (hi) atomicOr(&out32[(bp >> 5) + 1], hi);
            }
        }
    }
}

// Generic kernel for the rare chunks the fast kernel skips (codes longer than 16 bits, or a chunk
// too dense for the stage): sentinel table entries, one 64-bit insert per symbol, and for dense
// chunks straight atomicOr into the container.  Exits at once when there is no such chunk.
__global__ void __launch_bounds__(ENC_THREADS)
k_encode_slow(Ws ws, uint32_t *__restrict__ out32)
{
    __shared__ EncSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (ws.totals[TOT_ERR] || ws.totals[TOT_SLOW] == 0) return;
    const uint32_t nChunks = ws.totals[TOT_CHUNKS];
    const uint32_t per = (nChunks + gridDim.x - 1) / gridDim.x;
    const uint32_t cBeg = blockIdx.x * per, cEnd = min(nChunks, cBeg + per);
    const uint32_t tabLane = enc_smem_addr(sm.tab) + 4 * lane, stageA = enc_smem_addr(sm.stage);
    uint32_t curSeg = 0xFFFFFFFFu;
    for (uint32_t chunk = cBeg; chunk < cEnd; chunk++) {
        const uint32_t s = ws.chunkSeg[chunk];
        const uint32_t *si = ws.segInfo + (size_t)s * SI_WORDS;
        if (!ws.chunkSlow[chunk]) continue;             // the fast kernel's
        if (s != curSeg) {
            __syncthreads();
            uint32_t e = ws.segTab[(size_t)s * 256 + tid];      // ENC_THREADS == 256 symbols
            if (!e) e = 1;                              // absent symbol: zero-length sentinel (never looked up)
#pragma unroll 8
            for (int c = 0; c < 32; c++) sm.tab[tid * 32 + ((c + lane) & 31)] = e;
            curSeg = s;
            __syncthreads();
        }
        const uint32_t b = s >> 2, k = s & 3;
        const uint32_t c = chunk - si[SI_FIRST_CHUNK];
        const uint32_t S = si[SI_S];
        const uint32_t cnt = min((uint32_t)kChunkSyms, S - c * kChunkSyms);
        const uint32_t nRuns = cnt / ENC_RUN;
        const uint8_t *src = ws.strm[k] + (size_t)b * ws.strmStride[k] + (size_t)c * kChunkSyms;
        const unsigned long long gb0 = ws.segBitBase[s] + ws.chunkBitOff[chunk];
        // stage word k holds container word wBase + k; wBase is a multiple of 4, so 16-byte units
        // line up on both sides and the copy-out can use 128-bit loads and stores
        const uint32_t r0 = (uint32_t)(gb0 & 127);
        const unsigned long long wBase = (gb0 >> 5) & ~3ull;
        const bool active = (uint32_t)tid < nRuns;
        const bool dense = r0 + ws.chunkBits[chunk] > (uint32_t)ENC_STAGE_WORDS * 32;   // does not fit the stage

        // ---- pass 1: one lookup per symbol, entries stay in registers ------------------------
        uint32_t ent[32];
        uint32_t myBits = 0;
        if (active) {
            const uint4 a = reinterpret_cast<const uint4 *>(src)[tid * 2];
            const uint4 d = reinterpret_cast<const uint4 *>(src)[tid * 2 + 1];
            const uint32_t sy[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint32_t sym = (sy[j >> 2] >> ((j & 3) * 8)) & 0xFF;
                ent[j] = enc_lds(tabLane + sym * 128);
            }
#pragma unroll
            for (int j = 0; j < 32; j++) myBits += (uint32_t)(31 - __clz(ent[j]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) ent[j] = 1u;
        }
        uint32_t inc = myBits;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) sm.warpBits[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0, fullBits = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { uint32_t v = sm.warpBits[w]; if (w < warp) wbase += v; fullBits += v; }
        const uint32_t p0 = r0 + wbase + inc - myBits;   // staging bit position of this run

        if (!dense) {
            // ---- pass 2: pack into the stage; head word = first flushed word of the run --------
            uint32_t tail = 0, tailIdx = 0;
            if (active) {
                enc_pack_generic(ent, p0, stageA, tail, tailIdx);
            }
            // the tail of run t-1 lives in the low bits of the head word of run t
            const uint32_t prevTail = __shfl_up_sync(0xFFFFFFFFu, tail, 1);
            if (lane == 31) sm.warpTail[warp] = tail;
            if (active) {
                if (lane > 0) sm.stage[p0 >> 5] |= prevTail;                 // own word, own earlier store
                if ((uint32_t)tid == nRuns - 1) sm.stage[tailIdx] = tail;    // nobody follows: park the tail
            }
            __syncthreads();
            if (active && lane == 0 && warp > 0) sm.stage[p0 >> 5] |= sm.warpTail[warp - 1];
            __syncthreads();
            // ---- copy out.  Words [first,last) are complete; `first` may share its low bits with the
            // previous chunk / header bytes and `last` (if partial) its high bits with what follows:
            // those two are merged with atomicOr, whole 16-byte units go out as 128-bit stores.
            const uint32_t first = r0 >> 5, fbit = r0 & 31;
            const uint32_t endBit = r0 + fullBits, last = endBit >> 5, rem = endBit & 31;
            uint32_t *dst = out32 + wBase;
            const uint32_t vBeg = (first + (fbit ? 1u : 0u) + 3u) >> 2, vEnd = last >> 2;
            if (vBeg < vEnd) {
                for (uint32_t v = vBeg + tid; v < vEnd; v += ENC_THREADS)
                    reinterpret_cast<uint4 *>(dst)[v] = reinterpret_cast<const uint4 *>(sm.stage)[v];
                if (tid < 8) {                                                   // scalar edges
                    const uint32_t k = tid < 4 ? first + tid : 4 * vEnd + (tid - 4);
                    const bool inHead = tid < 4 && k < 4 * vBeg, inTail = tid >= 4 && k < last;
                    if (inHead || inTail) {
                        if (k == first && fbit) atomicOr(&dst[k], sm.stage[k]); else dst[k] = sm.stage[k];
                    }
                }
            } else {
                for (uint32_t k = first + tid; k < last; k += ENC_THREADS) {
                    if (k == first && fbit) atomicOr(&dst[k], sm.stage[k]); else dst[k] = sm.stage[k];
                }
            }
            if (tid == 0 && rem && nRuns) atomicOr(&dst[last], sm.stage[last]);
        } else if (active) {
            // ---- dense chunk (> 16 bits per symbol on average; adversarial): straight to the container
            unsigned long long bp = wBase * 32 + p0;
#pragma unroll 1
            for (int j = 0; j < 32; j++) {
                uint32_t e = 0;
#pragma unroll
                for (int q = 0; q < 32; q++) if (q == j) e = ent[q];
                const uint32_t len = (uint32_t)(31 - __clz(e));
                const uint32_t code = e ^ (1u << len);
                const unsigned long long v = (unsigned long long)code << (bp & 31);
                if ((uint32_t)v) atomicOr(&out32[bp >> 5], (uint32_t)v);
                if ((uint32_t)(v >> 32)) atomicOr(&out32[(bp >> 5) + 1], (uint32_t)(v >> 32));
                bp += len;
            }
        }
        // ---- ragged end of the stream (< 32 symbols): serial append ---------------------------
        if (tid == 0 && cnt > nRuns * ENC_RUN) {
            unsigned long long bp = gb0 + fullBits;
            for (uint32_t i = nRuns * ENC_RUN; i < cnt; i++) {
                const uint32_t e = sm.tab[(uint32_t)src[i] * 32];
                const uint32_t len = (uint32_t)(31 - __clz(e));
                const uint32_t code = e ^ (1u << len);
                const unsigned long long v = (unsigned long long)code << (bp & 31);
                if ((uint32_t)v) atomicOr(&out32[bp >> 5], (uint32_t)v);
                if ((uint32_t)(v >> 32)) atomicOr(&out32[(bp >> 5) + 1], (uint32_t)(v >> 32));
                bp += len;
            }
        }
        // no barrier here: the scan barrier of the next chunk orders this copy-out before the next
        // pack, and warpBits / warpTail are rewritten only after barriers every thread has passed
    }
}

// =========================================================================================
// headers: file header, record sizes, counters, serialised trees, raw bitmap byte
// =========================================================================================
__global__ void __launch_bounds__(128)
k_headers(Ws ws, uint8_t *__restrict__ out, uint32_t nBlocks, uint32_t flags,
          uint64_t hdrTotalBytes, uint32_t hdrBlocks, uint64_t batchStartCursor)
{
    if (ws.totals[TOT_ERR]) return;
    const int lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nBlocks) return;
    if (b == 0 && (flags & FCX_WITH_HEADER) && lane == 0) {
        uint8_t *h = out + batchStartCursor;              // mc:101-109 packed header, rewritten at mc:4128-4129
        h[0] = 'F'; h[1] = 'C'; h[2] = 'X';
        if (flags & FCX_FORMAT_X9) {                      // opt-in format: u64 byte count, u32 block count
            h[3] = '9';
            st_u32_bytes(h + 4, (uint32_t)hdrTotalBytes); st_u32_bytes(h + 8, (uint32_t)(hdrTotalBytes >> 32));
            st_u32_bytes(h + 12, hdrBlocks);
        } else {
            h[3] = '7';
            st_u32_bytes(h + 4, (uint32_t)hdrTotalBytes);     // uInt32 cmp_before_bytes (wraps)
            h[8] = (uint8_t)hdrBlocks; h[9] = (uint8_t)(hdrBlocks >> 8);   // uInt16 block_num (wraps)
        }
    }
    const unsigned long long off = ws.blkOff[b];
    const uint32_t *m = ws.meta + (size_t)b * M_WORDS;
    uint8_t *q = out + off;
    // payload size: next record offset is not known per block without a second array; recompute
    uint32_t sec[4];
    for (int k = 0; k < 4; k++) sec[k] = section_bytes(ws.segInfo + ((size_t)b * 4 + k) * SI_WORDS);
    const uint32_t payload = 4 + sec[0] + sec[1] + 4 + sec[2] + 4 + sec[3];
    if (lane == 0) { st_u32_bytes(q, payload); st_u32_bytes(q + 4, m[M_T]); }       // mc:4111, mc:2137-2139
    q += 8;
    for (int k = 0; k < 4; k++) {
        const uint32_t s = b * 4 + k;
        const uint32_t *si = ws.segInfo + (size_t)s * SI_WORDS;
        if (k == 2) { if (lane == 0) st_u32_bytes(q, m[M_PC]); q += 4; }             // mc:2186-2188
        if (k == 3) { if (lane == 0) st_u32_bytes(q, m[M_G]); q += 4; }              // mc:2225-2227
        const uint32_t S = si[SI_S];
        if (S == 0) continue;
        if (si[SI_RAW]) {                                                            // mc:2104-2107
            const uint8_t *src = ws.strm[k] + (size_t)b * ws.strmStride[k];
            for (uint32_t i = lane; i < S; i += 32) q[i] = src[i];
        } else {
            const uint32_t hb = si[SI_HDR];
            const uint8_t *h = ws.segHdr + (size_t)s * kSegHdrStride;
            for (uint32_t i = lane; i < hb; i += 32) q[i] = h[i];
            if (lane == 0) st_u32_bytes(q + hb, si[SI_W]);                           // mc:1081
        }
        q += sec[k];
    }
}

// =========================================================================================
// launches
// =========================================================================================
void huff_init()
{
    cudaFuncSetAttribute(k_encode_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncWSmem));
}

int launch_huff_stage(const Ws &ws, uint32_t nBlocks, uint8_t *out, uint64_t outCap, uint32_t flags,
                      uint64_t hdrTotalBytes, uint32_t hdrBlocks, uint64_t batchStartCursor, int numSM,
                      cudaStream_t st, void (*mark)(void *, int), void *markArg)
{
    const uint32_t nSeg = nBlocks * 4;
    int launches = 0;
    cudaMemsetAsync(ws.segHist, 0, (size_t)nSeg * 256 * sizeof(uint32_t), st);
    k_plan<<<1, 1024, 0, st>>>(ws, nBlocks); launches++;
    k_fill_chunks<<<(nSeg * 32 + 255) / 256, 256, 0, st>>>(ws, nSeg); launches++;
    if (mark) mark(markArg, 2);
    k_hist<<<numSM * 8, 256, 0, st>>>(ws); launches++;
    if (mark) mark(markArg, 3);
    k_tree<<<nSeg, 256, 0, st>>>(ws, nSeg, (flags & FCX_FORMAT_X9) ? 1u : 0u); launches++;
    if (mark) mark(markArg, 4);
    k_chunk_bits<<<numSM * 4, 256, 0, st>>>(ws); launches++;
    k_block_layout<<<1, 1024, 0, st>>>(ws, nBlocks, (flags & FCX_WITH_HEADER) ? ((flags & FCX_FORMAT_X9) ? FCX_X9_HEADER_BYTES : FCX_HEADER_BYTES) : 0u, outCap); launches++;
    // k_encode_w writes a sub-chunk that is too dense for its stage itself; only long-code streams go to k_encode_slow
    k_seg_scan<<<(nSeg + 7) / 8, 256, 0, st>>>(ws, nSeg); launches++;
    const uint32_t encGrid = (uint32_t)std::min(numSM * EW_CTAS, (int)kMaxEncCtas);
    k_enc_split<<<1, 1024, 0, st>>>(ws, nSeg, encGrid); launches++;
    if (mark) mark(markArg, 5);
    k_encode_w<<<encGrid, EW_THREADS, sizeof(EncWSmem), st>>>(ws, reinterpret_cast<uint32_t *>(out));
    launches++;
    if (mark) mark(markArg, 6);
    k_encode_slow<<<numSM * 2, ENC_THREADS, 0, st>>>(ws, reinterpret_cast<uint32_t *>(out)); launches++;
    if (mark) mark(markArg, 7);
    k_headers<<<(nBlocks + 3) / 4, 128, 0, st>>>(ws, out, nBlocks, flags, hdrTotalBytes, hdrBlocks, batchStartCursor);
    launches++;
    if (mark) mark(markArg, 8);
    return launches;
}

} // namespace fcx
