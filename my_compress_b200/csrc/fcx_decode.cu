// fcx_decode.cu -- LZ77 block decompressor (SURVEY 8f "next" row 1).
//
// Replaces my_decompress_file_lz77 (mc:2255-2393): my_huffman_decode_char (mc:1107-1187) /
// huffman_decode_char (mc:930-984) for the four sections, decombine_bits (mc:1315-1338),
// golomb_rice_decode (mc:309-358) and my_LZ77_decompress (mc:1716-1735).  Output is what the
// reference produces, including a single-symbol stream decoding to zeros
// and the "Fatal Error" early stop when matches run out (mc:2337-2340).
//
// The format has no resynchronisation points, so a Huffman section and the Golomb stream can
// only be read front to back; the parallelism is across blocks (one CTA each, ~7 CTAs per SM,
// so ~1000 blocks decode at once) and across the four sections of a block (one warp each):
//   * every warp builds an 11-bit look-up table of its tree in shared memory (an entry holds
//     up to three symbols whose codes fit the window) and decodes warp-uniformly: all lanes
//     carry the same bit buffer, code words are fetched 32 at a time with one coalesced load and
//     handed out by shuffle, a slow bit-by-bit walk takes over for long codes and at the ends;
//   * the warp of the Golomb section then parses the l values the same way (count the ones
//     with one ffs, two remainder bits);
//   * warp 0 replays the tokens 32 at a time: flag bits -> match ranks (popc), a warp scan of
//     len+1 gives every token its output position, tokens whose source lies before the batch copy
//     in parallel inside a 16 KB shared ring of recent output, the few that reach into the
//     batch follow in order (overlapping copies use source offset k mod p), and finished words
//     are flushed to HBM with coalesced stores.
// Malformed payloads that would send the reference out of bounds yield length 0xFFFFFFFF;
// a match length above 257 (the encoder's maximum) or a distance of 0 is treated as malformed.
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

constexpr int DEC_LUT_BITS = 11;
constexpr int DEC_LUT = 1 << DEC_LUT_BITS;
constexpr uint32_t DEC_RING = 16384;                              // >= 2047 (window) + 32 * 258 (one batch)
constexpr uint32_t DEC_LUT_INVALID = 0xFFFFFFFFu;

struct DecSmem {
    union {
        uint32_t lut[4][DEC_LUT];        // phase 1: bits 0-3 code bits used, 4-5 symbols, 8.. up to three symbols
        uint8_t ring[DEC_RING];          // phase 2: the most recent output bytes, addressed by position & (DEC_RING-1)
    };
    uint8_t pair[4][512];
    uint8_t map[4][64];
    uint32_t secOff[4], secCnt[4], secTn[4], secW[4], secRaw[4];
    uint32_t T, Pc, G, bad, nL;
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

// ---- warp-uniform reader of little-endian 32-bit words at any byte address --------------------
// Every lane holds the same state; words come 32 at a time (lane i loads word base+i, the next
// 32 are requested while the current ones are consumed) and are broadcast by shuffle.
struct BitReader {
    const uint32_t *aligned;     // base rounded down to 4 bytes
    uint32_t sh;                 // (address & 3) * 8
    uint32_t W;                  // words in the stream
    uint32_t wi;                 // next word to hand out
    uint32_t curW, nextW;        // this lane's word of the current / next group of 32
    unsigned long long buf;
    uint32_t nbits;

    __device__ __forceinline__ uint32_t load_word(uint32_t i) const
    {
        if (i >= W) return 0u;
        const uint32_t a0 = __ldcg(aligned + i);
        if (sh == 0) return a0;
        return __funnelshift_r(a0, __ldcg(aligned + i + 1), sh);
    }
    __device__ __forceinline__ void init(const uint8_t *p, uint32_t words)
    {
        const int lane = threadIdx.x & 31;
        aligned = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
        sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
        W = words; wi = 0; buf = 0; nbits = 0;
        curW = load_word(lane);
        nextW = load_word(32 + lane);
    }
    // appends one word to the bit buffer (call only when nbits <= 32); false when the stream is exhausted
    __device__ __forceinline__ bool refill()
    {
        if (wi >= W) return false;
        const uint32_t w = __shfl_sync(0xFFFFFFFFu, curW, wi & 31);
        buf |= (unsigned long long)w << nbits;
        nbits += 32;
        wi++;
        if ((wi & 31) == 0) {
            curW = nextW;
            nextW = load_word(wi + 32 + (threadIdx.x & 31));
        }
        return true;
    }
};

// packed byte writer: symbols are collected four at a time and stored as one word by lane 0
struct ByteWriter {
    uint32_t *dst32;
    unsigned long long acc;
    uint32_t have;               // bytes waiting in acc
    uint32_t j;                  // bytes produced so far
    __device__ __forceinline__ void init(uint8_t *dst) { dst32 = reinterpret_cast<uint32_t *>(dst); acc = 0; have = 0; j = 0; }
    __device__ __forceinline__ void put(uint32_t syms, uint32_t n)
    {
        acc |= (unsigned long long)syms << (8 * have);
        have += n; j += n;
        if (have >= 4) {
            if ((threadIdx.x & 31) == 0) dst32[(j - have) >> 2] = (uint32_t)acc;
            acc >>= 32; have -= 4;
        }
    }
    __device__ __forceinline__ void finish()      // buffers are zero-filled and padded: a whole word is fine
    {
        if (have && (threadIdx.x & 31) == 0) dst32[(j - have) >> 2] = (uint32_t)acc;
    }
};

// One Huffman section, decoded by a whole warp (uniformly).  Returns false on a malformed tree.
__device__ bool decode_section(DecSmem &sm, int k, const uint8_t *pl, uint8_t *dst)
{
    const int lane = threadIdx.x & 31;
    const uint32_t at = sm.secOff[k], cnt = sm.secCnt[k], Tn = sm.secTn[k], W = sm.secW[k];
    if (sm.secRaw[k]) {
        for (uint32_t i = lane; i < cnt; i += 32) dst[i] = pl[at + i];
        return true;
    }
    // Tn == 0 (single-symbol stream): nothing is decoded, the zero-filled buffer stands (App. B1)
    if (!(cnt && Tn && W)) return true;
    const uint32_t mapB = (2 * Tn + 7) / 8;
    for (uint32_t i = lane; i < mapB; i += 32) sm.map[k][i] = pl[at + 1 + i];
    for (uint32_t i = lane; i < 2 * Tn; i += 32) sm.pair[k][i] = pl[at + 1 + mapB + i];
    __syncwarp();
    const uint32_t R = Tn + 1, root = Tn - 1;                        // root = last stored node (mc:945-946)
    const uint8_t *pairK = sm.pair[k], *mapK = sm.map[k];
    // ---- look-up table: greedy decode of the 11 index bits, up to three symbols ---------------
    for (uint32_t idx = lane; idx < (uint32_t)DEC_LUT; idx += 32) {
        uint32_t node = root, used = 0, ns = 0, syms = 0, e = 0;
        for (int bit = 0; bit < DEC_LUT_BITS; bit++) {
            const uint32_t c = 2 * node + ((idx >> bit) & 1u);
            const uint32_t ci = pairK[c];
            if (!((mapK[c >> 3] >> (c & 7)) & 1)) {                  // leaf (mc:968)
                syms |= ci << (8 * ns);
                ns++; used = bit + 1; node = root;
                if (ns == 3) break;
            } else {
                node = ci - (256 - R);                               // mc:976-977
                if (node >= Tn) { if (ns == 0) e = DEC_LUT_INVALID; break; }
            }
        }
        if (e != DEC_LUT_INVALID) e = ns ? (used | (ns << 4) | (syms << 8)) : 0u;
        sm.lut[k][idx] = e;
    }
    __syncwarp();
    const uint32_t *lut = sm.lut[k];
    BitReader br;
    br.init(pl + at + 1 + mapB + 2 * Tn + 4, W);
    ByteWriter bw;
    bw.init(dst);
    bool bad = false;
    while (bw.j < cnt) {
        if (br.nbits <= 32) br.refill();
        const uint32_t e = lut[(uint32_t)br.buf & (DEC_LUT - 1)];
        const uint32_t used = e & 15u, ns = (e >> 4) & 3u;
        if (e != 0 && e != DEC_LUT_INVALID && used <= br.nbits && bw.j + ns <= cnt) {
            bw.put(e >> 8, ns);
            br.buf >>= used; br.nbits -= used;
            continue;
        }
        // ---- exact path: one symbol, bit by bit (long codes, the last bits, the last symbols) ----
        uint32_t node = root;
        bool done = false;
        for (;;) {
            if (br.nbits == 0 && !br.refill()) { done = true; break; }        // out of code words (mc:955)
            const uint32_t c = 2 * node + ((uint32_t)br.buf & 1u);
            br.buf >>= 1; br.nbits--;
            const uint32_t ci = pairK[c];
            if (!((mapK[c >> 3] >> (c & 7)) & 1)) { bw.put(ci, 1); break; }
            node = ci - (256 - R);
            if (node >= Tn) { bad = true; done = true; break; }
        }
        if (done) break;
    }
    bw.finish();
    return !bad;
}

// golomb_rice_decode (mc:309-358) for up to Pc values: l = 4*q + r, q ones, a zero, two remainder bits.
// Stops at the first value whose bits are not all there (the reference's overrun check); returns how
// many values were stored.
__device__ uint32_t parse_golomb(const uint8_t *gb, uint32_t G, uint32_t Pc, uint16_t *lv)
{
    const int lane = threadIdx.x & 31;
    BitReader br;
    br.init(gb, G);
    uint32_t m = 0;
    while (m < Pc) {
        uint32_t q = 0;
        bool ok = true;
        for (;;) {                                                   // the run of ones
            if (br.nbits <= 32 && !br.refill() && br.nbits == 0) { ok = false; break; }
            const unsigned long long inv = ~br.buf;
            const uint32_t ones = inv ? (uint32_t)(__ffsll((long long)inv) - 1) : 64u;
            if (ones >= br.nbits) { q += br.nbits; br.buf = 0; br.nbits = 0; continue; }   // run goes on in the next word
            q += ones;
            br.buf >>= ones; br.nbits -= ones;
            break;
        }
        if (!ok) break;
        if (br.nbits < 3) br.refill();
        if (br.nbits < 3) break;                                     // gpos + 3 > 32*G
        const uint32_t r = ((uint32_t)br.buf >> 1) & 3u;
        br.buf >>= 3; br.nbits -= 3;
        const uint32_t l = min(q * 4u + r, 0xFFFFu);
        if (lane == 0) lv[m] = (uint16_t)l;
        m++;
    }
    return m;
}

__global__ void __launch_bounds__(128)
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
    uint8_t *o = out + (size_t)b * (kBlockBytes + 8);

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

    // ---- phase 1: the four sections, one warp each; the Golomb warp goes on to the l values --------
    {
        const bool ok = decode_section(sm, warp, pl, dst[warp]);
        if (!ok && lane == 0) sm.bad = 1;
        if (warp == 3) {
            __syncwarp();
            __threadfence_block();                                   // lane 0's stores of the Golomb bytes, read back below
            const uint32_t nL = ok ? parse_golomb(dst[3], sm.G, sm.Pc, lv) : 0u;
            if (lane == 0) sm.nL = nL;
        }
    }
    __syncthreads();
    if (sm.bad) { if (tid == 0) outLen[b] = DEC_BAD; return; }
    if (warp != 0) return;

    // ---- phase 2: token replay, 32 tokens per trip (my_LZ77_decompress, mc:1716-1735) -----------------
    const uint32_t T = sm.T, Pc = sm.Pc, nL = sm.nL;
    const uint32_t *bm32 = reinterpret_cast<const uint32_t *>(dst[0]);
    const uint8_t *cs = dst[1], *pk = dst[2];
    uint8_t *ring = sm.ring;
    uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
    constexpr uint32_t RM = DEC_RING - 1;
    uint32_t cur = 0, mBase = 0, flushed = 0;
    bool bad = false, stop = false;
    const uint32_t ltMask = (1u << lane) - 1u;
    for (uint32_t t0 = 0; t0 < T && !stop; t0 += 32) {
        const uint32_t t = t0 + lane;
        bool valid = t < T;
        const uint32_t word = __ldcg(bm32 + (t0 >> 5));
        bool isMatch = valid && !((word >> lane) & 1u);
        const uint32_t mmask = __ballot_sync(0xFFFFFFFFu, isMatch);
        const uint32_t mIdx = mBase + __popc(mmask & ltMask);
        const uint32_t stopMask = __ballot_sync(0xFFFFFFFFu, isMatch && mIdx >= Pc);
        if (stopMask) {                                              // mc:2337-2340: "Fatal Error", stop before that token
            const int first = __ffs(stopMask) - 1;
            valid = valid && lane < first;
            isMatch = isMatch && valid;
            stop = true;
        }
        uint32_t l = 0, p = 0;
        bool badLane = false;
        if (isMatch) {
            if (mIdx >= nL) badLane = true;                          // its Golomb code ran past the stream
            else {
                l = __ldcg(lv + mIdx);
                const uint32_t bp = kPBits * mIdx;                   // decombine_bits, mc:1315-1338
                const uint32_t by = bp >> 3;
                const uint32_t raw = (uint32_t)__ldcg(pk + by) | ((uint32_t)__ldcg(pk + by + 1) << 8) | ((uint32_t)__ldcg(pk + by + 2) << 16);
                p = (raw >> (bp & 7)) & 0x7FFu;
                if (l > (uint32_t)kMaxLen) badLane = true;
            }
        }
        const uint32_t len1 = valid ? l + 1 : 0;
        uint32_t inc = len1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += v;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        const uint32_t pos = cur + inc - len1;
        if (valid && !badLane && ((l && (p == 0 || p > pos)) || (unsigned long long)pos + l + 1 > kBlockBytes + 8)) badLane = true;
        if (__any_sync(0xFFFFFFFFu, badLane)) { bad = true; break; }
        const uint32_t lit = valid ? (uint32_t)__ldcg(cs + t) : 0u;
        // a token is independent when all its source bytes were produced before this trip
        const bool dep = isMatch && l && (pos - p + l > cur);
        if (valid) {
            if (!dep) for (uint32_t k = 0; k < l; k++) ring[(pos + k) & RM] = ring[(pos - p + k) & RM];   // mc:1730-1732
            ring[(pos + l) & RM] = (uint8_t)lit;                                                         // mc:1734
        }
        __syncwarp();
        uint32_t depMask = __ballot_sync(0xFFFFFFFFu, dep);
        while (depMask) {                                            // in order: each may read what the previous wrote
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
        // ---- flush the finished whole words to HBM ---------------------------------------------------
        const uint32_t newCur = cur + total;
        const uint32_t fe = newCur & ~3u;
        for (uint32_t w = (flushed >> 2) + lane; w < (fe >> 2); w += 32)
            o32[w] = reinterpret_cast<const uint32_t *>(ring)[w & (RM >> 2)];
        flushed = fe;
        cur = newCur;
        mBase += __popc(mmask);
        __syncwarp();
    }
    if (!bad) for (uint32_t i = flushed + lane; i < cur; i += 32) o[i] = ring[i & RM];
    if (lane == 0) outLen[b] = bad ? DEC_BAD : cur;
}

int launch_decode(const uint8_t *d_file, const uint64_t *d_recOff, uint32_t nBlocks, uint8_t *d_out,
                  uint32_t *d_outLen, uint8_t *d_scratch, size_t scratchPerBlock, cudaStream_t st)
{
    // the reference decodes into zero-initialised buffers (mc:2270,2282,2291,2305)
    cudaMemsetAsync(d_scratch, 0, scratchPerBlock * nBlocks, st);
    k_decode<<<nBlocks, 128, 0, st>>>(d_file, d_recOff, d_out, d_outLen, d_scratch, scratchPerBlock);
    return 1;
}

} // namespace fcx
