// fcx_parse.cu -- greedy parse + stream split (subsystem 2 of the north star).
//
// Replaces my_LZ77_compress (mc:1675-1714: cursor += l+1), the literal-flag bitmap of
// make_bitMap_table (mc:2073-2094), the c/p/l split (mc:2150-2161), combine_bits at 11 bits
// (mc:1292-1313) and golomb_rice_encode (mc:258-304).  Input is the per-position (len, dist)
// table of k_match; output is the four byte streams the Huffman stage encodes.
//
// The parse is a dependent chain next(i) = i + len(i) + 1 with steps of 1..258.  It runs as ONE kernel over
// tiles of 4096 positions (k_parse), every tile doing three things:
//   exit     for each of the 258 offsets a chain can enter the tile at, the offset at which it enters the next
//            one (the tile's exit function, from the jump tables below).  Chains that start at different offsets
//            fall into step after a few tokens, so the function is almost always CONSTANT: the tile then
//            publishes its exit at once, without knowing its own entry.  Otherwise it waits for the exit of the
//            tile before it (one word per tile, a chain of waits only through runs of such tiles -- e.g. a block
//            of one repeated byte, whose 258-byte tokens never merge) and publishes exit(entry).
//   counts   with its entry known the tile walks its part of the chain: tokens, matches, Golomb bits; it
//            publishes them and learns the rank of its first token from the counts of the tiles before it (one
//            64-bit word per tile, all read at once: the device-wide scan of the parse)
//   emit     c bytes, flag bits, 11-bit distances and Golomb codes are packed in shared memory during that wait
//            and copied to their final bit positions
// CTAs are numbered block-first (blockIdx.x = block, blockIdx.y = tile): the CTAs resident at any time are the
// same few tiles of MANY blocks, so the tile a CTA waits for was dispatched a whole row of CTAs earlier and has
// normally finished; a CTA only ever waits for lower-numbered ones.
// Inside a tile the chain is resolved by a three-level jump table:
//   J1[i] = first chain position at or beyond the end of i's 16-position mini-tile
//   J2[i] = first chain position at or beyond the end of i's 256-position mid-tile
// both built by backward sweeps (16 dependent steps each, all threads busy); 16 hops then
// give the entry of every mid-tile, 16 more the entry of every mini-tile, and each thread
// walks its own mini-tile (<= 16 tokens).  A mini-tile without a match (most of them on incompressible
// input) skips its sweep and its walks: its jumps, counts, flag bits and c bytes are known from its entry.
#include "fcx_dev.cuh"

namespace fcx {

constexpr int PT = kParseTile;
constexpr int PT_THREADS = 256;
constexpr int PT_LOOK = 272;
constexpr int PT_JPAD = PT + 2 * (PT / 16);
constexpr int PT_ENTRIES = kParseEntries;         // 258 possible entry offsets

struct JumpSmem {
    uint32_t lc[PT / 4];
    uint16_t J1[PT_JPAD];
    uint16_t J2[PT_JPAD];
    uint16_t midEntry[16];
    uint16_t miniEntry[256];
    uint32_t warpTok[8], warpGb[8];
};

__device__ __forceinline__ int jphys(int i) { return i + 2 * (i >> 4); }
__device__ __forceinline__ int code_len(uint32_t code) { return code ? (int)code + 2 : 0; }

// All-literal tiles (incompressible input: ~60 % of the tiles of random data) need no jump tables: every
// position from the entry on is a one-byte token, so the exit offset is 0, the counts are (limit - entry, 0, 0)
// and the emit is a byte copy plus a run of set flag bits.  CTA-uniform test on the tile's len codes.
__device__ __forceinline__ bool tile_has_match(const uint8_t *lcG, int tb)
{
    const uint4 v = reinterpret_cast<const uint4 *>(lcG + tb)[threadIdx.x];
    return __syncthreads_or((v.x | v.y | v.z | v.w) != 0u) != 0;
}

// stage the tile's len codes and build J1 / J2 (ends with a __syncthreads).  Returns whether this thread's own
// mini-tile (positions 16 tid .. 16 tid + 15) holds no match.
__device__ __forceinline__ bool build_jumps(JumpSmem &sm, const uint8_t *lcG, int tb, int limit)
{
    const int tid = threadIdx.x;
    const uint4 v = reinterpret_cast<const uint4 *>(lcG + tb)[tid];
    reinterpret_cast<uint4 *>(sm.lc)[tid] = v;
    const int base = tid * 16, pad = 2 * tid;              // jphys(i) = i + pad for every i of this thread's mini-tile
    // codes beyond the block's end are stale (k_match writes none there): they count as literals
    const bool plain = (v.x | v.y | v.z | v.w) == 0u || base >= limit;
    if (plain) {
        // (a') nothing but literals: every position leads to the end of the mini-tile
        uint32_t *w = reinterpret_cast<uint32_t *>(sm.J1) + 9 * tid;     // (base + pad) / 2: 18 tid is even
        const uint32_t two = (uint32_t)(base + 16) * 0x10001u;
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = two;
    } else {   // (a) mini-tile backward sweep
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 15; j >= 0; j--) {
            const int i = base + j;
            const uint32_t code = (w[j >> 2] >> ((j & 3) * 8)) & 0xFF;
            const int len = (i < limit) ? code_len(code) : 0;
            const int nx = i + len + 1;
            const uint16_t val = (nx >= base + 16) ? (uint16_t)nx : sm.J1[nx + pad];    // nx stays in this mini-tile
            sm.J1[i + pad] = val;
        }
    }
    __syncthreads();
    {   // (b) mid-tile backward sweep: 16 lanes per mid-tile, lockstep
        const int M = tid >> 4, j = tid & 15;
        const int midEnd = (M + 1) * 256;
        for (int m = 15; m >= 0; m--) {
            const int ip = M * 256 + m * 16 + j + 2 * (M * 16 + m);         // jphys of position M*256 + m*16 + j
            const int a = sm.J1[ip];
            const uint16_t val = (a >= midEnd) ? (uint16_t)a : sm.J2[jphys(a)];
            sm.J2[ip] = val;
            __syncwarp();
        }
    }
    __syncthreads();
    return plain;
}

// entries of the 16 mid-tiles and 256 mini-tiles for a known tile entry (ends with a __syncthreads)
__device__ __forceinline__ void resolve_entries(JumpSmem &sm, int entry)
{
    const int tid = threadIdx.x;
    if (tid == 0) {
        int e = entry;
        for (int M = 0; M < 16; M++) {
            sm.midEntry[M] = (uint16_t)e;
            if (e < (M + 1) * 256) e = sm.J2[jphys(e)];
        }
    }
    __syncthreads();
    if (tid < 16) {
        const int M = tid;
        int e = sm.midEntry[M];
        for (int m = 0; m < 16; m++) {
            sm.miniEntry[M * 16 + m] = (uint16_t)e;
            if (e < M * 256 + (m + 1) * 16) e = sm.J1[jphys(e)];
        }
    }
    __syncthreads();
}

// walk of one mini-tile: counts
__device__ __forceinline__ void walk_count(const uint8_t *lcb, int e0, int hi, uint32_t &tok, uint32_t &mat, uint32_t &gb)
{
    tok = 0; mat = 0; gb = 0;
    for (int e = e0; e < hi;) {
        const int len = code_len(lcb[e]);
        tok++;
        if (len) { mat++; gb += (uint32_t)(len >> 2) + 3; }
        e += len + 1;
    }
}

// ---- 3: emit -----------------------------------------------------------------------------------------
struct EmitSmem {
    JumpSmem j;
    uint32_t xs[(PT + PT_LOOK) / 4];
    uint8_t cbuf[PT + 16];
    uint32_t fbits[PT / 32 + 4];          // tile-relative bit streams: bit 0 = the tile's first flag / distance / Golomb bit
    uint32_t pbits[368];
    uint32_t gbits[128];
    unsigned long long base;              // look-back result: packed exclusive ranks of the tile's first token
    uint32_t exit0, entry;
};

// One status word per tile (ws.tileState, zeroed per launch): ready << 62 | Golomb bits << 40 | matches << 21 | tokens.
// Sums over a block stay inside the fields (tokens <= 2^20, matches <= 2^18, Golomb bits < 2^20), so packed words add
// without carries between fields.
constexpr unsigned long long TS_MASK = (1ull << 62) - 1;
__device__ __forceinline__ unsigned long long ts_pack(uint32_t tok, uint32_t mat, uint32_t gb)
{
    return (unsigned long long)tok | ((unsigned long long)mat << 21) | ((unsigned long long)gb << 40);
}
__device__ __forceinline__ uint32_t ts_tok(unsigned long long v) { return (uint32_t)v & 0x1FFFFFu; }
__device__ __forceinline__ uint32_t ts_mat(unsigned long long v) { return (uint32_t)(v >> 21) & 0x7FFFFu; }
__device__ __forceinline__ uint32_t ts_gb(unsigned long long v) { return (uint32_t)(v >> 40) & 0x1FFFFFu; }
__device__ __forceinline__ void ts_store(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ts_load(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One exit word per tile (ws.tileExit, zeroed per launch): bit 31 = known, low bits = the offset at which the
// block's chain enters the NEXT tile.
__device__ __forceinline__ void exit_publish(uint32_t *row, int tile, uint32_t exitOff)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(row + tile), "r"(0x80000000u | exitOff) : "memory");
}
__device__ __forceinline__ uint32_t exit_wait(const uint32_t *row, int tile)
{
    uint32_t v;
    do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(row + tile) : "memory");
    } while (!(v & 0x80000000u));
    return v & 0x7FFFFFFFu;
}

// Ranks by one warp (all 32 lanes): sum of the counts of tiles 0..tile-1 of this block.  Every tile publishes its own
// counts as soon as it knows them (before it packs its tokens), so by the time a tile asks, the words of the tiles
// before it are usually all there: lane j reads words j, j + 32, ... (at most eight, all loads in flight at once) and
// only spins on the ones not yet written.  A CTA only ever waits for lower-numbered ones (earlier tiles of its
// block), which were dispatched before it.
__device__ __forceinline__ unsigned long long tile_lookback(const unsigned long long *row, int tile, int lane)
{
    static_assert(kParseTilesPerBlock == 256, "eight words per lane");
    unsigned long long v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { const int j = lane + 32 * k; v[k] = j < tile ? ts_load(row + j) : (1ull << 62); }
    unsigned long long acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int j = lane + 32 * k;
        while ((v[k] >> 62) == 0) v[k] = ts_load(row + j);
        if (j < tile) acc += v[k] & TS_MASK;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    return acc;
}

// waits for the ranks of the tile's first token (its own counts were published earlier by publish_counts); the last
// tile of a block also writes the block's meta words.  Called by warp 0 (all lanes); returns the packed exclusive ranks.
__device__ __forceinline__ void publish_counts(const Ws &ws, int b, int tile, unsigned long long mine)
{
    ts_store(ws.tileState + (size_t)b * kParseTilesPerBlock + tile, (1ull << 62) | mine);
}
__device__ __forceinline__ unsigned long long tile_ranks(const Ws &ws, int b, int tile, int nTiles, int lane, unsigned long long mine)
{
    const unsigned long long *row = ws.tileState + (size_t)b * kParseTilesPerBlock;
    const unsigned long long before = tile_lookback(row, tile, lane);
    if (lane == 0 && tile == nTiles - 1) {
        const unsigned long long incl = before + mine;
        const uint32_t ta = ts_tok(incl), tc = ts_mat(incl), tg = ts_gb(incl);
        const uint32_t G = (tg + 31) >> 5;                        // mc:295-302: last partial word is flushed
        uint32_t *m = ws.meta + (size_t)b * M_WORDS;
        m[M_T] = ta; m[M_PC] = tc; m[M_G] = G; m[M_GBITS] = tg;
        m[M_S0] = (ta + 7) >> 3;                                  // mc:2077-2078
        m[M_S1] = ta;
        m[M_S2] = (kPBits * tc) / 8 + 1;                          // mc:2193: one spare byte when divisible
        m[M_S3] = 4 * G;                                          // mc:2230-2236: words viewed as bytes
    }
    return before;
}

// OR `nbits` (<= 32) low bits of v into a zero-initialised shared bit buffer at bit `pos`
__device__ __forceinline__ void or_bits(uint32_t *buf, uint32_t pos, uint32_t v, int nbits)
{
    if (nbits <= 0) return;
    const uint32_t w = pos >> 5, sh = pos & 31;
    atomicOr(&buf[w], v << sh);
    if (sh + nbits > 32) atomicOr(&buf[w + 1], v >> (32 - sh));
}

// copy a tile's bit buffer (bit 0 = its first bit; zero beyond nbits, one spare word) into a pre-zeroed global bit
// stream at absolute bit position gpos: the words are shifted into place on the way
__device__ __forceinline__ void flush_bits(uint32_t *gstream, const uint32_t *buf, unsigned long long gpos, uint32_t nbits)
{
    if (nbits == 0) return;
    const uint32_t off = (uint32_t)(gpos & 31);
    const unsigned long long w0 = gpos >> 5;
    const uint32_t total = off + nbits, nWords = (total + 31) >> 5, nSrc = (nbits + 31) >> 5;
    for (uint32_t j = threadIdx.x; j < nWords; j += PT_THREADS) {
        const uint32_t lo = j ? buf[j - 1] : 0u, hi = j < nSrc ? buf[j] : 0u;
        const uint32_t v = __funnelshift_l(lo, hi, off);              // (hi << off) | (lo >> (32 - off)); off = 0: hi
        const bool partial = (j == 0 && off) || (j == nWords - 1 && (total & 31));
        if (partial) atomicOr(&gstream[w0 + j], v);                   // word shared with a neighbouring tile
        else gstream[w0 + j] = v;
    }
}

__global__ void __launch_bounds__(PT_THREADS)
k_parse(const uint8_t *__restrict__ in, const uint32_t *__restrict__ blkLen,
        const uint8_t *__restrict__ lenCode, const uint16_t *__restrict__ dist, Ws ws)
{
    __shared__ EmitSmem sm;
    const int b = blockIdx.x, tile = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;   // block-first numbering
    const int n = (int)blkLen[b];
    const int tb = tile * PT;
    if (tb >= n) return;
    const int nTiles = (n + PT - 1) / PT;
    const int limit = min(PT, n - tb);
    const uint8_t *x = in + (size_t)b * kBlockBytes;
    const uint16_t *dsG = dist + (size_t)b * kBlockBytes;
    uint32_t *exitRow = ws.tileExit + (size_t)b * kParseTilesPerBlock;
    if (!tile_has_match(lenCode + (size_t)b * kBlockBytes, tb)) {
        // every token is a literal: the chain leaves at offset 0 whatever its entry, c = the bytes themselves,
        // flags = a run of ones at the tile's token rank
        if (tid == 0) exit_publish(exitRow, tile, 0u);
        // the tile's bytes travel to shared memory while thread 0 asks the tile before this one for the entry
        {
            const int pos = tb + tid * 16;
            uint4 q;
            if (pos + 16 <= n) q = *reinterpret_cast<const uint4 *>(x + pos);
            else {
                uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                for (int k = 0; k < 16; k++)
                    if (pos + k < n) w[k >> 2] |= (uint32_t)x[pos + k] << ((k & 3) * 8);
                q = make_uint4(w[0], w[1], w[2], w[3]);
            }
            reinterpret_cast<uint4 *>(sm.xs)[tid] = q;
        }
        if (warp == 0) {
            uint32_t e = 0;
            if (lane == 0 && tile) e = exit_wait(exitRow, tile - 1);
            e = __shfl_sync(0xFFFFFFFFu, e, 0);
            const uint32_t cnt = (uint32_t)max(limit - (int)e, 0);
            if (lane == 0) { sm.entry = e; publish_counts(ws, b, tile, ts_pack(cnt, 0, 0)); }
            const unsigned long long r = tile_ranks(ws, b, tile, nTiles, lane, ts_pack(cnt, 0, 0));
            if (lane == 0) sm.base = r;
        }
        __syncthreads();
        const int entry = (int)sm.entry;
        const uint32_t tileTok = (uint32_t)max(limit - entry, 0);
        const uint32_t baseTok = ts_tok(sm.base);
        const uint8_t *src = reinterpret_cast<const uint8_t *>(sm.xs) + entry;
        uint8_t *cOut = ws.strm[1] + (size_t)b * ws.strmStride[1] + baseTok;
        // byte copy between two unaligned ranges: 4-byte stores on the destination's alignment, bytes at the edges
        const uint32_t head = min(tileTok, (uint32_t)((4u - ((uint32_t)(size_t)cOut & 3u)) & 3u));
        if ((uint32_t)tid < head) cOut[tid] = src[tid];
        const uint32_t nW = (tileTok - head) >> 2;
        for (uint32_t k = tid; k < nW; k += PT_THREADS)
            reinterpret_cast<uint32_t *>(cOut + head)[k] = ld32u(sm.xs, (int)(entry + head + 4 * k));
        const uint32_t done = head + 4 * nW;
        if ((uint32_t)tid < tileTok - done) cOut[done + tid] = src[done + tid];
        // flag bits [baseTok, baseTok + tileTok) of the (pre-zeroed) bitmap stream
        if (tileTok) {
            uint32_t *fs = reinterpret_cast<uint32_t *>(ws.strm[0] + (size_t)b * ws.strmStride[0]);
            const uint32_t b0 = baseTok, b1 = baseTok + tileTok;             // bit range
            const uint32_t w0 = b0 >> 5, w1 = (b1 - 1) >> 5;
            for (uint32_t w = w0 + tid; w <= w1; w += PT_THREADS) {
                uint32_t m = 0xFFFFFFFFu;
                if (w == w0) m &= 0xFFFFFFFFu << (b0 & 31);
                if (w == w1 && (b1 & 31)) m &= 0xFFFFFFFFu >> (32 - (b1 & 31));
                if (m == 0xFFFFFFFFu) fs[w] = m; else atomicOr(&fs[w], m);   // edge words are shared with the neighbouring tiles
            }
        }
        return;
    }
    // stage bytes (+ look-ahead for c = x[i+len]) and zero the bit buffers; distances are read
    // straight from global memory, only for the tokens that are matches
    for (int v = tid; v < (PT + PT_LOOK) / 16; v += PT_THREADS) {
        const int pos = tb + v * 16;
        uint4 q;
        if (pos + 16 <= n) q = *reinterpret_cast<const uint4 *>(x + pos);
        else {
            uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (pos + k < n) w[k >> 2] |= (uint32_t)x[pos + k] << ((k & 3) * 8);
            q = make_uint4(w[0], w[1], w[2], w[3]);
        }
        reinterpret_cast<uint4 *>(sm.xs)[v] = q;
    }
    for (int k = tid; k < PT / 32 + 4; k += PT_THREADS) sm.fbits[k] = 0;
    for (int k = tid; k < 368; k += PT_THREADS) sm.pbits[k] = 0;
    for (int k = tid; k < 128; k += PT_THREADS) sm.gbits[k] = 0;
    const bool plain = build_jumps(sm.j, lenCode + (size_t)b * kBlockBytes, tb, limit);

    // ---- exit function: where does a chain entering at offset e leave?  All 258 chains advance one mid-tile per
    // round (one J2 hop) until they stand on the same position -- they merge after a token or two, so that is
    // round 1 or 2 -- and from there ONE thread follows the common chain to the end of the tile.  If they never
    // meet the exit depends on the entry.
    {
        int e = tid, e2 = PT_THREADS + tid;                    // entries 0..255; threads 0, 1 also carry entries 256, 257
        bool constant = false;
        for (int lvl = 1; lvl <= PT / 256; lvl++) {
            const int bound = lvl * 256;
            if (e < bound) e = sm.j.J2[jphys(e)];              // leaves mid-tile lvl - 1 (a long match may already be beyond)
            if (tid < PT_ENTRIES - PT_THREADS && e2 < bound) e2 = sm.j.J2[jphys(e2)];
            if (tid == 0) sm.exit0 = (uint32_t)e;
            __syncthreads();
            const bool same = e == (int)sm.exit0 && (tid >= PT_ENTRIES - PT_THREADS || e2 == (int)sm.exit0);
            if (__syncthreads_and(same)) { constant = true; break; }
        }
        // entry: the exit of the tile before this one.  A constant tile tells its successor before it knows its own.
        if (tid == 0) {
            if (constant) {
                int ex = e;
                while (ex < PT) ex = sm.j.J2[jphys(ex)];
                exit_publish(exitRow, tile, (uint32_t)(ex - PT));
            }
            const uint32_t en = tile ? exit_wait(exitRow, tile - 1) : 0u;
            if (!constant) {
                int ex = (int)en;
                while (ex < PT) ex = sm.j.J2[jphys(ex)];
                exit_publish(exitRow, tile, (uint32_t)(ex - PT));
            }
            sm.entry = en;
        }
        __syncthreads();
    }
    resolve_entries(sm.j, (int)sm.entry);

    const uint8_t *lcb = reinterpret_cast<const uint8_t *>(sm.j.lc);
    const uint8_t *xsb = reinterpret_cast<const uint8_t *>(sm.xs);
    const int hi = min(tid * 16 + 16, limit);
    const int e0 = sm.j.miniEntry[tid];
    uint32_t myTok, myMat, myGb;
    if (plain) { myTok = (uint32_t)max(hi - e0, 0); myMat = 0; myGb = 0; }      // literals only: one token per byte from the entry on
    else walk_count(lcb, e0, hi, myTok, myMat, myGb);
    // CTA exclusive scan of (tok | mat<<16) and gb
    const uint32_t packed = myTok | (myMat << 16);
    uint32_t incP = packed, incG = myGb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t tp = __shfl_up_sync(0xFFFFFFFFu, incP, d), tg = __shfl_up_sync(0xFFFFFFFFu, incG, d);
        if (lane >= d) { incP += tp; incG += tg; }
    }
    if (lane == 31) { sm.j.warpTok[warp] = incP; sm.j.warpGb[warp] = incG; }
    __syncthreads();
    uint32_t baseP = 0, baseG = 0, tileP = 0, tileG = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint32_t p = sm.j.warpTok[w], g = sm.j.warpGb[w];
        if (w < warp) { baseP += p; baseG += g; }
        tileP += p; tileG += g;
    }
    const uint32_t exP = baseP + incP - packed;
    uint32_t tokIdx = exP & 0xFFFF, matIdx = exP >> 16, gOff = baseG + incG - myGb;
    const uint32_t tileTok = tileP & 0xFFFF, tileMat = tileP >> 16;
    // the tile's counts are known: tell the tiles behind this one before the packing starts
    const unsigned long long mine = ts_pack(tileTok, tileMat, tileG);
    if (tid == 0) publish_counts(ws, b, tile, mine);

    // A thread's tokens have consecutive ranks, so its flag bits (<= 16) and packed distances
    // (<= 5 x 11 bits) are contiguous: gather them in registers and merge once, instead of one
    // shared-memory atomic per token.  Bit positions are relative to the tile's first token.
    const uint32_t tok0 = tokIdx, mat0 = matIdx;
    uint32_t flagReg = 0;
    unsigned long long pReg = 0;
    if (plain) {
        for (uint32_t k = 0; k < myTok; k++) sm.cbuf[tok0 + k] = xsb[e0 + (int)k];
        flagReg = (1u << myTok) - 1u;                                             // myTok <= 16
    } else {
        for (int e = e0; e < hi;) {
            const int len = code_len(lcb[e]);
            sm.cbuf[tokIdx] = xsb[e + len];                                           // mc:1489: byte after the match
            if (len == 0) {
                flagReg |= 1u << (tokIdx - tok0);                                     // mc:2087-2089: bit set = literal
            } else {
                pReg |= (unsigned long long)dsG[tb + e] << (kPBits * (matIdx - mat0));  // mc:1301-1309
                const uint32_t gp = gOff;
                const int q = len >> 2;                                               // mc:267-269: q ones, 0, 2 remainder bits
                if (q <= 29) {                                                        // whole code fits 32 bits
                    or_bits(sm.gbits, gp, ((1u << q) - 1u) | ((uint32_t)(len & 3) << (q + 1)), q + 3);
                } else {
                    if (q > 32) { or_bits(sm.gbits, gp, 0xFFFFFFFFu, 32); or_bits(sm.gbits, gp + 32, 0xFFFFFFFFu >> (64 - q), q - 32); }
                    else or_bits(sm.gbits, gp, 0xFFFFFFFFu >> (32 - q), q);
                    or_bits(sm.gbits, gp + q + 1, (uint32_t)(len & 3), 2);
                }
                matIdx++;
                gOff += (uint32_t)q + 3;
            }
            tokIdx++;
            e += len + 1;
        }
    }
    if (flagReg) or_bits(sm.fbits, tok0, flagReg, 16);
    if (matIdx > mat0) {
        const uint32_t pp = kPBits * mat0;
        or_bits(sm.pbits, pp, (uint32_t)pReg, 32);
        if (matIdx - mat0 > 2) or_bits(sm.pbits, pp + 32, (uint32_t)(pReg >> 32), 32);
    }
    // ranks of the tile's first token: by now the tiles before this one have usually published theirs
    if (warp == 0) {
        const unsigned long long r = tile_ranks(ws, b, tile, nTiles, lane, mine);
        if (lane == 0) sm.base = r;
    }
    __syncthreads();
    const uint32_t baseTok = ts_tok(sm.base), baseMat = ts_mat(sm.base), baseGb = ts_gb(sm.base);
    uint8_t *cOut = ws.strm[1] + (size_t)b * ws.strmStride[1] + baseTok;
    for (uint32_t k = tid; k < tileTok; k += PT_THREADS) cOut[k] = sm.cbuf[k];
    flush_bits(reinterpret_cast<uint32_t *>(ws.strm[0] + (size_t)b * ws.strmStride[0]), sm.fbits, baseTok, tileTok);
    flush_bits(reinterpret_cast<uint32_t *>(ws.strm[2] + (size_t)b * ws.strmStride[2]), sm.pbits,
               (unsigned long long)kPBits * baseMat, kPBits * tileMat);
    flush_bits(reinterpret_cast<uint32_t *>(ws.strm[3] + (size_t)b * ws.strmStride[3]), sm.gbits, baseGb, tileG);
}

void parse_init() {}

int launch_parse(const uint8_t *in, const uint32_t *blkLen, const Ws &ws, uint32_t nBlocks, uint32_t maxLen, cudaStream_t st)
{
    const dim3 grid(nBlocks, (maxLen + PT - 1) / PT);               // block-first: see the file comment
    // flag / distance / Golomb streams are assembled with OR: start from zero (also yields the
    // spare zero byte of the p stream and the zero padding of the last Golomb word)
    cudaMemsetAsync(ws.strm[0], 0, (size_t)nBlocks * ws.strmStride[0], st);
    cudaMemsetAsync(ws.strm[2], 0, (size_t)nBlocks * ws.strmStride[2], st);
    cudaMemsetAsync(ws.strm[3], 0, (size_t)nBlocks * ws.strmStride[3], st);
    cudaMemsetAsync(ws.tileState, 0, (size_t)nBlocks * kParseTilesPerBlock * sizeof(unsigned long long), st);
    cudaMemsetAsync(ws.tileExit, 0, (size_t)nBlocks * kParseTilesPerBlock * sizeof(uint32_t), st);
    k_parse<<<grid, PT_THREADS, 0, st>>>(in, blkLen, ws.lenCode, ws.dist, ws);
    return 1;
}

} // namespace fcx
