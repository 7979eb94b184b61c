// fcx_decode.cu -- LZ77 block decompressor (SURVEY 8f "next" row 1).
//
// Replaces my_decompress_file_lz77 (mc:2255-2393): my_huffman_decode_char (mc:1107-1187) /
// huffman_decode_char (mc:930-984) for the four sections, decombine_bits (mc:1315-1338),
// golomb_rice_decode (mc:309-358) and my_LZ77_decompress (mc:1716-1735).  Output is what the
// reference produces, including a single-symbol stream decoding to zeros
// and the "Fatal Error" early stop when matches run out (mc:2337-2340).
//
// The format has no resynchronisation points, so the sequential dependency of a prefix-coded
// stream is broken speculatively (see spec_decode): one CTA per block, its 128 lanes decode
// the four Huffman sections one after the other (11-bit look-up table in shared memory, tree
// walk for longer codes), then the Golomb stream of l values with the same scheme.  All lanes then
// gather 1024 tokens at a time into shared memory (flag bits -> match ranks by popc and a scan)
// and warp 0 replays them 32 per trip: a warp scan of len+1 gives every token its output
// position, tokens whose source lies before the batch copy in parallel
// inside a 16 KB shared ring of recent output, the few that reach into the batch follow in
// order (overlapping copies use source offset k mod p), and finished words are flushed to HBM
// with coalesced stores.
// Malformed payloads that would send the reference out of bounds yield length 0xFFFFFFFF;
// a match length above 257 (the encoder's maximum), a distance of 0 or a block that would decode to more
// than 1 MiB (the reference tolerates 8 bytes more) is treated as malformed.
#include "fcx_dev.cuh"

namespace fcx {

constexpr uint32_t DEC_BAD = 0xFFFFFFFFu;
constexpr size_t DEC_OFF_BM = 0;
constexpr size_t DEC_OFF_CS = 131072 + 256;
constexpr size_t DEC_OFF_PK = DEC_OFF_CS + 1048576 + 256;
constexpr size_t DEC_OFF_GB = DEC_OFF_PK + 360704 + 256;
constexpr size_t DEC_OFF_LV = DEC_OFF_GB + 105216 + 256;          // u16 l value per match
constexpr uint32_t DEC_MAX_PC = kBlockBytes / 4 + 1;
constexpr size_t DEC_SCRATCH = DEC_OFF_LV + 2 * (size_t)DEC_MAX_PC + 254;
static_assert(DEC_SCRATCH % 16 == 0 && DEC_OFF_LV % 16 == 0, "scratch regions stay 16-byte aligned");

size_t decode_scratch_per_block() { return DEC_SCRATCH; }

constexpr int DEC_THREADS = 128;
constexpr int DEC_LUT_BITS = 11;
constexpr int DEC_LUT = 1 << DEC_LUT_BITS;
constexpr uint32_t DEC_RING = 16384;                              // >= 2047 (window) + 32 * 258 (one batch)
constexpr uint32_t DEC_SUB_BITS = 512;                            // bits a lane decodes per trip
constexpr uint32_t DEC_FIN = 0xFFFFFFFFu;                         // "the stream ended inside this subsequence"
constexpr int DEC_SUPER = 8 * DEC_THREADS;                        // tokens gathered per trip
constexpr uint32_t DEC_TOK_NONE = 0xFFFD, DEC_TOK_STOP = 0xFFFE, DEC_TOK_BAD = 0xFFFF;

struct DecSmem {
    uint8_t ring[DEC_RING];              // replay: the most recent output bytes, addressed by position & (DEC_RING-1)
    uint16_t lut[DEC_LUT];               // current section: len << 8 | symbol for codes of <= 11 bits, else 0
    uint8_t pair[512];
    uint8_t map[64];
    uint32_t endS[DEC_THREADS];          // bit position at which each lane's subsequence ends (DEC_FIN: stream over)
    uint32_t warpTot[DEC_THREADS / 32];
    uint32_t secOff[4], secCnt[4], secTn[4], secW[4], secRaw[4];
    uint32_t T, Pc, G, bad, nL, fin;
    uint16_t tokL[DEC_SUPER];            // replay: match length per token (0 literal-only), or a DEC_TOK_* marker
    uint16_t tokP[DEC_SUPER];            //         distance
    uint8_t tokC[DEC_SUPER];             //         literal byte
};

__device__ __forceinline__ uint32_t rd32(const uint8_t *p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Parses one section header at payload offset `at`; returns the offset just past the section.
__device__ uint32_t parse_section(DecSmem &sm, int k, const uint8_t *pl, uint32_t plen, uint32_t at, uint32_t cnt)
{
    sm.secCnt[k] = cnt; sm.secRaw[k] = 0; sm.secOff[k] = at; sm.secTn[k] = 0; sm.secW[k] = 0;
    if (at + 5 > plen) { sm.bad = 1; return at; }
    const uint32_t Tn = pl[at];
    const uint32_t mapB = (2 * Tn + 7) / 8;
    const uint32_t wAt = at + 1 + mapB + 2 * Tn;
    if (wAt + 4 > plen) { sm.bad = 1; return at; }
    const uint32_t W = rd32(pl + wAt);
    if ((unsigned long long)wAt + 4 + 4ull * W > plen) { sm.bad = 1; return at; }
    sm.secTn[k] = Tn; sm.secW[k] = W;
    return wAt + 4 + 4 * W;
}

// ---- per-lane bit cursor over little-endian 32-bit words at any byte address -------------------
template <bool kConst>
struct LaneBits {
    const uint32_t *a;           // base rounded down to 4 bytes
    uint32_t sh, W;              // (address & 3) * 8, words in the stream
    uint32_t wi, nbits, pos;     // next word, valid bits in buf, bit index of buf's bit 0
    unsigned long long buf;
    __device__ __forceinline__ uint32_t ld(const uint32_t *p) const { return kConst ? __ldg(p) : __ldcg(p); }
    __device__ __forceinline__ uint32_t word(uint32_t i) const
    {
        const uint32_t w0 = ld(a + i);
        return sh ? __funnelshift_r(w0, ld(a + i + 1), sh) : w0;
    }
    __device__ __forceinline__ void open(const uint8_t *p, uint32_t words)
    {
        a = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
        sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
        W = words; wi = 0; nbits = 0; pos = 0; buf = 0;
    }
    __device__ __forceinline__ void seek(uint32_t bit)
    {
        pos = bit; wi = bit >> 5; buf = 0; nbits = 0;
        if (wi < W) { buf = word(wi) >> (bit & 31); nbits = 32 - (bit & 31); wi++; }
    }
    __device__ __forceinline__ void fill()
    {
        if (nbits <= 32 && wi < W) { buf |= (unsigned long long)word(wi) << nbits; nbits += 32; wi++; }
    }
    __device__ __forceinline__ void skip(uint32_t n) { buf >>= n; nbits -= n; pos += n; }       // n < 64
    __device__ __forceinline__ void drop_all() { pos += nbits; buf = 0; nbits = 0; }
};

// next(): 0 = one symbol decoded, 1 = the stream ended first, 2 = malformed tree
struct HuffCodec {
    const uint16_t *lut;
    const uint8_t *pairK, *mapK;
    uint32_t Tn, R, root;
    template <class B>
    __device__ __forceinline__ int next(B &lb, uint32_t &sym) const
    {
        lb.fill();
        if (lb.nbits == 0) return 1;                                 // out of code words (mc:955)
        const uint32_t e = lut[(uint32_t)lb.buf & (DEC_LUT - 1)];
        const uint32_t len = e >> 8;
        if (e && len <= lb.nbits) { sym = e & 0xFFu; lb.skip(len); return 0; }
        uint32_t node = root;                                        // long code, or the last bits: walk the tree
        for (;;) {
            if (lb.nbits == 0) { lb.fill(); if (lb.nbits == 0) return 1; }
            const uint32_t c = 2 * node + ((uint32_t)lb.buf & 1u);
            lb.skip(1);
            const uint32_t ci = pairK[c];
            if (!((mapK[c >> 3] >> (c & 7)) & 1)) { sym = ci; return 0; }   // leaf (mc:968)
            node = ci - (256 - R);                                   // mc:976-977
            if (node >= Tn) return 2;
        }
    }
};

// golomb_rice_decode (mc:309-358): q ones, a zero, two remainder bits -> l = 4*q + r.  A value whose
// bits are not all there ends the stream (the reference's overrun check).
struct GolombCodec {
    template <class B>
    __device__ __forceinline__ int next(B &lb, uint32_t &sym) const
    {
        uint32_t q = 0;
        for (;;) {
            lb.fill();
            if (lb.nbits == 0) return 1;
            const unsigned long long inv = ~lb.buf;
            const uint32_t ones = inv ? (uint32_t)(__ffsll((long long)inv) - 1) : 64u;
            if (ones >= lb.nbits) { q += lb.nbits; lb.drop_all(); continue; }   // the run goes on in the next word
            q += ones; lb.skip(ones);
            break;
        }
        lb.fill();
        if (lb.nbits < 3) return 1;                                  // gpos + 3 > 32*G
        sym = min(4u * q + (((uint32_t)lb.buf >> 1) & 3u), 0xFFFFu);
        lb.skip(3);
        return 0;
    }
};

// Speculative parallel decode of one prefix-coded stream by the whole CTA.  Every lane takes the
// next DEC_SUB_BITS bits; lane 0 starts on a code boundary, the others guess.  A lane decodes until
// it crosses into the next lane's bits and publishes where it ended; a lane whose predecessor ended
// somewhere else than where it started decodes again from there, until nobody moves (prefix codes
// fall into step after a few symbols, so this is one or two rounds; by induction over the lanes the
// fixed point is the sequential decode whatever the data).  A scan of the symbol counts then gives
// every lane its output offset for the writing pass.  Returns the symbols produced (<= cnt).
template <class Codec, class Bits, class Emit>
__device__ uint32_t spec_decode(DecSmem &sm, const Codec &cd, Bits lb, uint32_t cnt, Emit emit, bool &bad)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t totalBits = lb.W * 32u;
    uint32_t segStart = 0, outBase = 0;
    while (outBase < cnt && segStart < totalBits) {                  // CTA-uniform
        const uint32_t bndNext = segStart + (uint32_t)(tid + 1) * DEC_SUB_BITS;
        uint32_t myStart = segStart + (uint32_t)tid * DEC_SUB_BITS, myEnd = DEC_FIN, myN = 0;
        bool myBad = false;
        auto run = [&]() {
            myN = 0; myBad = false;
            if (myStart == DEC_FIN || myStart >= totalBits) { myEnd = DEC_FIN; return; }
            lb.seek(myStart);
            uint32_t sym;
            while (lb.pos < bndNext) {
                const int st = cd.next(lb, sym);
                if (st) { myBad = st == 2; myEnd = DEC_FIN; return; }
                myN++;
            }
            myEnd = lb.pos;
        };
        run();
        sm.endS[tid] = myEnd;
        for (;;) {
            __syncthreads();
            const uint32_t prevEnd = tid ? sm.endS[tid - 1] : segStart;
            const bool moved = prevEnd != myStart;
            if (!__syncthreads_or(moved)) break;                     // also: everyone has read before anyone rewrites
            if (moved) { myStart = prevEnd; run(); sm.endS[tid] = myEnd; }
        }
        // ---- output offsets ------------------------------------------------------------------------
        uint32_t inc = myN;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += v;
        }
        if (lane == 31) sm.warpTot[warp] = inc;
        __syncthreads();
        uint32_t off = inc - myN, total = 0;
#pragma unroll
        for (int w = 0; w < DEC_THREADS / 32; w++) { const uint32_t v = sm.warpTot[w]; if (w < warp) off += v; total += v; }
        const uint32_t first = outBase + off;
        // the sequential decoder would walk into this lane's malformed code unless cnt stops it earlier
        if (myBad && first + myN < cnt) bad = true;
        // ---- writing pass: the same symbols again, now with a place to go ------------------------------
        if (myN && first < cnt) {
            const uint32_t lim = min(myN, cnt - first);
            lb.seek(myStart);
            uint32_t sym;
            for (uint32_t i = 0; i < lim; i++) {
                if (cd.next(lb, sym)) break;
                emit(first + i, sym);
            }
        }
        segStart = sm.endS[DEC_THREADS - 1];
        outBase += total;
        __syncthreads();                                             // endS / warpTot are rewritten by the next trip
        if (segStart == DEC_FIN) break;
    }
    return min(outBase, cnt);
}

__global__ void __launch_bounds__(DEC_THREADS)
k_decode(const uint8_t *__restrict__ file, const uint64_t *__restrict__ recOff, uint8_t *__restrict__ out,
         uint32_t *__restrict__ outLen, uint8_t *__restrict__ scratchAll, size_t scratchPerBlock)
{
    __shared__ DecSmem sm;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *rec = file + recOff[b];
    const uint32_t plen = rd32(rec);
    const uint8_t *pl = rec + 4;
    uint8_t *scr = scratchAll + (size_t)b * scratchPerBlock;
    uint8_t *dst[4] = {scr + DEC_OFF_BM, scr + DEC_OFF_CS, scr + DEC_OFF_PK, scr + DEC_OFF_GB};
    uint16_t *lv = reinterpret_cast<uint16_t *>(scr + DEC_OFF_LV);
    uint8_t *o = out + (size_t)b * kBlockBytes;                    // blocks sit back to back: a full block is 1 MiB

    if (tid == 0) {
        sm.bad = 0; sm.Pc = 0; sm.G = 0; sm.T = 0; sm.nL = 0;
        if (plen < 4) sm.bad = 1;
        else {
            const uint32_t T = rd32(pl);
            sm.T = T;
            if (T > kBlockBytes) sm.bad = 1;
            else {
                uint32_t at = 4;
                const uint32_t mapB = (T + 7) / 8;
                if (mapB > 1) at = parse_section(sm, 0, pl, plen, at, mapB);            // mc:2273-2276
                else { sm.secRaw[0] = 1; sm.secOff[0] = at; sm.secCnt[0] = mapB; sm.secW[0] = 0; sm.secTn[0] = 0;
                       if (at + mapB > plen) sm.bad = 1; at += mapB; }                  // mc:2277-2280
                if (!sm.bad) at = parse_section(sm, 1, pl, plen, at, T);                // mc:2283
                if (!sm.bad && at + 4 <= plen) {
                    sm.Pc = rd32(pl + at); at += 4;                                     // mc:2286
                    if (sm.Pc > DEC_MAX_PC) sm.bad = 1;
                } else sm.bad = 1;
                if (!sm.bad) at = parse_section(sm, 2, pl, plen, at, (kPBits * sm.Pc) / 8 + 1);   // mc:2290-2292
                if (!sm.bad && at + 4 <= plen) {
                    sm.G = rd32(pl + at); at += 4;                                      // mc:2302
                    if (sm.G > 105216 / 4) sm.bad = 1;
                } else sm.bad = 1;
                if (!sm.bad) {
                    if (sm.G) at = parse_section(sm, 3, pl, plen, at, sm.G * 4);        // mc:2308
                    else { sm.secCnt[3] = 0; sm.secW[3] = 0; sm.secTn[3] = 0; sm.secRaw[3] = 0; sm.secOff[3] = at; }   // App. B8
                }
            }
        }
    }
    __syncthreads();
    if (sm.bad) { if (tid == 0) outLen[b] = DEC_BAD; return; }

    // ---- phase 1: the four sections one after the other, all lanes on each; then the l values ------
    for (int k = 0; k < 4; k++) {
        __syncthreads();                                             // the previous section's table is no longer in use
        const uint32_t at = sm.secOff[k], cnt = sm.secCnt[k], Tn = sm.secTn[k], W = sm.secW[k];
        if (sm.secRaw[k]) {
            for (uint32_t i = tid; i < cnt; i += DEC_THREADS) dst[k][i] = pl[at + i];
            continue;
        }
        // Tn == 0 (single-symbol stream): nothing is decoded, the zero-filled buffer stands (App. B1)
        if (!(cnt && Tn && W)) continue;
        const uint32_t mapB = (2 * Tn + 7) / 8;
        for (uint32_t i = tid; i < mapB; i += DEC_THREADS) sm.map[i] = pl[at + 1 + i];
        for (uint32_t i = tid; i < 2 * Tn; i += DEC_THREADS) sm.pair[i] = pl[at + 1 + mapB + i];
        __syncthreads();
        const uint32_t R = Tn + 1, root = Tn - 1;                    // root = last stored node (mc:945-946)
        for (uint32_t idx = tid; idx < (uint32_t)DEC_LUT; idx += DEC_THREADS) {
            uint32_t node = root, e = 0;
            for (int bit = 0; bit < DEC_LUT_BITS; bit++) {
                const uint32_t c = 2 * node + ((idx >> bit) & 1u);
                const uint32_t ci = sm.pair[c];
                if (!((sm.map[c >> 3] >> (c & 7)) & 1)) { e = ((uint32_t)(bit + 1) << 8) | ci; break; }
                node = ci - (256 - R);
                if (node >= Tn) break;                               // malformed: the exact walk reports it if it is ever reached
            }
            sm.lut[idx] = (uint16_t)e;
        }
        __syncthreads();
        HuffCodec cd{sm.lut, sm.pair, sm.map, Tn, R, root};
        LaneBits<true> lb;
        lb.open(pl + at + 1 + mapB + 2 * Tn + 4, W);
        uint8_t *d = dst[k];
        bool bad = false;
        spec_decode(sm, cd, lb, cnt, [d](uint32_t i, uint32_t sym) { d[i] = (uint8_t)sym; }, bad);
        if (bad) sm.bad = 1;
    }
    __syncthreads();
    if (sm.bad) { if (tid == 0) outLen[b] = DEC_BAD; return; }
    {
        uint32_t nL = 0;
        if (sm.G && sm.Pc) {
            GolombCodec gc;
            LaneBits<false> gbits;
            gbits.open(dst[3], sm.G);
            bool bad = false;
            nL = spec_decode(sm, gc, gbits, sm.Pc, [lv](uint32_t i, uint32_t sym) { lv[i] = (uint16_t)sym; }, bad);
        }
        if (tid == 0) sm.nL = nL;
    }
    __syncthreads();

    // ---- phase 2: token replay (my_LZ77_decompress, mc:1716-1735) -------------------------------------
    // 1024 tokens at a time: all lanes gather (len, distance, literal) of eight tokens each into shared
    // memory (flag bits -> match ranks by popc and a CTA scan, then independent loads), warp 0 replays
    // them 32 per trip.
    const uint32_t T = sm.T, Pc = sm.Pc, nL = sm.nL;
    const uint8_t *bm = dst[0], *cs = dst[1], *pk = dst[2];
    uint8_t *ring = sm.ring;
    uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
    constexpr uint32_t RM = DEC_RING - 1;
    uint32_t cur = 0, mBase = 0, flushed = 0;                        // cur / flushed live in warp 0
    bool bad = false;
    if (tid == 0) sm.fin = 0;
    for (uint32_t tS = 0; tS < T; tS += DEC_SUPER) {
        // ---- gather -------------------------------------------------------------------------------------
        const uint32_t tb = tS + 8u * tid;
        const uint32_t nvalid = tb < T ? min(8u, T - tb) : 0u;
        const uint32_t flags = nvalid ? (uint32_t)__ldcg(bm + (tb >> 3)) : 0xFFu;
        const uint32_t zmask = ~flags & ((1u << nvalid) - 1u);       // match tokens among mine
        uint32_t inc = (uint32_t)__popc(zmask);
        const uint32_t mine = inc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += v;
        }
        if (lane == 31) sm.warpTot[warp] = inc;
        __syncthreads();
        uint32_t rank = mBase + inc - mine, tot = 0;
#pragma unroll
        for (int w = 0; w < DEC_THREADS / 32; w++) { const uint32_t v = sm.warpTot[w]; if (w < warp) rank += v; tot += v; }
        mBase += tot;
        unsigned long long lits = 0;
        if (nvalid) lits = __ldcg(reinterpret_cast<const unsigned long long *>(cs + tb));   // 8-byte aligned, buffer is padded
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t L = DEC_TOK_NONE, P = 0;
            if ((uint32_t)i < nvalid) {
                L = 0;
                if ((zmask >> i) & 1u) {
                    const uint32_t m = rank + (uint32_t)__popc(zmask & ((1u << i) - 1u));
                    if (m >= Pc) L = DEC_TOK_STOP;                   // mc:2337-2340: "Fatal Error", stop before this token
                    else if (m >= nL) L = DEC_TOK_BAD;               // its Golomb code ran past the stream
                    else {
                        L = __ldcg(lv + m);
                        const uint32_t bp = kPBits * m, by = bp >> 3;   // decombine_bits, mc:1315-1338
                        const uint32_t raw = (uint32_t)__ldcg(pk + by) | ((uint32_t)__ldcg(pk + by + 1) << 8) |
                                             ((uint32_t)__ldcg(pk + by + 2) << 16);
                        P = (raw >> (bp & 7)) & 0x7FFu;
                        if (L > (uint32_t)kMaxLen) L = DEC_TOK_BAD;
                    }
                }
            }
            sm.tokL[8 * tid + i] = (uint16_t)L;
            sm.tokP[8 * tid + i] = (uint16_t)P;
            sm.tokC[8 * tid + i] = (uint8_t)(lits >> (8 * i));
        }
        __syncthreads();
        // ---- replay ---------------------------------------------------------------------------------------
        if (warp == 0) {
            const uint32_t nTok = min((uint32_t)DEC_SUPER, T - tS);
            bool stop = false;
            for (uint32_t bb = 0; bb < nTok && !stop && !bad; bb += 32) {
                uint32_t l = sm.tokL[bb + lane];
                const uint32_t p = sm.tokP[bb + lane], lit = sm.tokC[bb + lane];
                bool valid = l != DEC_TOK_NONE;
                const uint32_t special = __ballot_sync(0xFFFFFFFFu, l == DEC_TOK_STOP || l == DEC_TOK_BAD);
                if (special) {
                    const int first = __ffs(special) - 1;
                    if (__shfl_sync(0xFFFFFFFFu, l, first) == DEC_TOK_BAD) { bad = true; break; }
                    valid = valid && lane < first;
                    stop = true;
                }
                if (!valid) l = 0;
                const uint32_t len1 = valid ? l + 1 : 0;
                uint32_t sc = len1;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, sc, d);
                    if (lane >= d) sc += v;
                }
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, sc, 31);
                const uint32_t pos = cur + sc - len1;
                const bool badLane = valid && ((l && (p == 0 || p > pos)) || (unsigned long long)pos + l + 1 > kBlockBytes);
                if (__any_sync(0xFFFFFFFFu, badLane)) { bad = true; break; }
                // a token is independent when all its source bytes were produced before this trip
                const bool dep = valid && l && (pos - p + l > cur);
                if (valid) {
                    if (!dep) for (uint32_t k = 0; k < l; k++) ring[(pos + k) & RM] = ring[(pos - p + k) & RM];   // mc:1730-1732
                    ring[(pos + l) & RM] = (uint8_t)lit;                                                         // mc:1734
                }
                __syncwarp();
                uint32_t depMask = __ballot_sync(0xFFFFFFFFu, dep);
                while (depMask) {                                    // in order: each may read what the previous wrote
                    const int d = __ffs(depMask) - 1;
                    depMask &= depMask - 1;
                    const uint32_t pd = __shfl_sync(0xFFFFFFFFu, pos, d), ld = __shfl_sync(0xFFFFFFFFu, l, d),
                                   sd = __shfl_sync(0xFFFFFFFFu, p, d);
                    // the source window [pd-sd, pd) is final; an overlapping copy repeats it with period sd
                    for (uint32_t k = lane; k < ld; k += 32) {
                        const uint32_t off = sd >= ld ? k : k % sd;
                        ring[(pd + k) & RM] = ring[(pd - sd + off) & RM];
                    }
                    __syncwarp();
                }
                // ---- flush the finished whole words to HBM ---------------------------------------------
                const uint32_t newCur = cur + total;
                const uint32_t fe = newCur & ~3u;
                for (uint32_t w = (flushed >> 2) + lane; w < (fe >> 2); w += 32)
                    o32[w] = reinterpret_cast<const uint32_t *>(ring)[w & (RM >> 2)];
                flushed = fe;
                cur = newCur;
                __syncwarp();
            }
            if (lane == 0 && (stop || bad)) sm.fin = 1;
        }
        __syncthreads();
        if (sm.fin) break;
    }
    if (warp == 0) {
        if (!bad) for (uint32_t i = flushed + lane; i < cur; i += 32) o[i] = ring[i & RM];
        if (lane == 0) outLen[b] = bad ? DEC_BAD : cur;
    }
}

void decode_init() {}

int launch_decode(const uint8_t *d_file, const uint64_t *d_recOff, uint32_t nBlocks, uint8_t *d_out,
                  uint32_t *d_outLen, uint8_t *d_scratch, size_t scratchPerBlock, cudaStream_t st)
{
    // the reference decodes into zero-initialised buffers (mc:2270,2282,2291,2305)
    cudaMemsetAsync(d_scratch, 0, scratchPerBlock * nBlocks, st);
    k_decode<<<nBlocks, DEC_THREADS, 0, st>>>(d_file, d_recOff, d_out, d_outLen, d_scratch, scratchPerBlock);
    return 1;
}

} // namespace fcx
