// fcx_decode.cu -- LZ77 block decompressor (SURVEY 8f "next" row 1).
//
// Replaces my_decompress_file_lz77 (mc:2255-2393): my_huffman_decode_char (mc:1107-1187) /
// huffman_decode_char (mc:930-984) for the four sections, decombine_bits (mc:1315-1338),
// golomb_rice_decode (mc:309-358) and my_LZ77_decompress (mc:1716-1735).  Output is what the
// reference produces, including a single-symbol stream decoding to zeros
// and the "Fatal Error" early stop when matches run out (mc:2337-2340).
//
// Three kernels per batch of blocks:
//   k_dec_parse     one thread per block walks the payload's section headers (offsets, symbol counts, tree
//                   sizes, word counts) into a small per-block record and rejects truncated payloads
//   k_dec_sections  one CTA per (section, block): the four Huffman sections of a block decode side by side.
//                   The format has no resynchronisation points, so the sequential dependency of a prefix-coded
//                   stream is broken speculatively (see spec_decode): every lane takes the next 512 bits, lane
//                   0 starts on a code boundary, the others guess, and lanes re-decode until nobody's start
//                   moves (prefix codes fall into step after a few symbols).  11-bit look-up table in shared
//                   memory, tree walk for longer codes.
//   k_dec_replay    one CTA per block: the Golomb stream of match lengths by the same speculative scheme, then
//                   the token replay, BYTE-parallel: for a batch of up to 2048 tokens / 8 KiB of output a scan
//                   of len+1 places every token, every output byte gets a pointer to the byte it copies
//                   (itself for a literal), pointer doubling over the batch collapses chains of copies -- a
//                   byte ends up pointing either before the batch or at a literal; at most 13 rounds whatever
//                   the data, one or two in practice because the encoder prefers the farthest match -- and all
//                   bytes are then fetched at once.  A 16 KiB shared ring holds the window and the batch;
//                   finished words leave for HBM with coalesced stores.
// Malformed payloads that would send the reference out of bounds yield length 0xFFFFFFFF;
// a match length above 257 (the encoder's maximum), a distance of 0 or a block that would decode to more
// than 1 MiB (the reference tolerates 8 bytes more) is treated as malformed.
#include "fcx_bits.cuh"

namespace fcx {

constexpr uint32_t DEC_BAD = 0xFFFFFFFFu;
constexpr size_t DEC_OFF_BM = 0;
constexpr size_t DEC_OFF_CS = 131072 + 256;
constexpr size_t DEC_OFF_PK = DEC_OFF_CS + 1048576 + 256;
constexpr size_t DEC_OFF_GB = DEC_OFF_PK + 360704 + 256;
constexpr size_t DEC_OFF_LV = DEC_OFF_GB + 105216 + 256;          // u16 l value per match
constexpr uint32_t DEC_MAX_PC = kBlockBytes / 4 + 1;
constexpr size_t DEC_OFF_META = DEC_OFF_LV + 2 * (size_t)DEC_MAX_PC + 254;
constexpr size_t DEC_SCRATCH = DEC_OFF_META + 256;
static_assert(DEC_SCRATCH % 16 == 0 && DEC_OFF_LV % 16 == 0 && DEC_OFF_META % 16 == 0, "scratch regions stay 16-byte aligned");

size_t decode_scratch_per_block() { return DEC_SCRATCH; }

// per-block record written by k_dec_parse (at DEC_OFF_META of the block's scratch)
struct DecMeta {
    uint32_t T, Pc, G, bad;
    uint32_t secOff[4], secCnt[4], secTn[4], secW[4], secRaw[4];
    uint32_t secSym[4];                // FCX9: 0x100 | symbol of a single-symbol section (0 otherwise)
    uint32_t nL;                       // l values the Golomb stream really held (k_dec_replay)
};
static_assert(sizeof(DecMeta) <= 256, "the record fits its scratch slot");

// =========================================================================================
// 1: section headers
// =========================================================================================
// Parses one section header at payload offset `at`; returns the offset just past the section.
// x9: the opt-in FCX9 format, where a section without tree nodes carries its one symbol behind the zero count.
__device__ uint32_t parse_section(DecMeta &m, int k, const uint8_t *pl, uint32_t plen, uint32_t at, uint32_t cnt, bool x9)
{
    m.secCnt[k] = cnt; m.secRaw[k] = 0; m.secOff[k] = at; m.secTn[k] = 0; m.secW[k] = 0; m.secSym[k] = 0;
    if (at + 5 > plen) { m.bad = 1; return at; }
    const uint32_t Tn = pl[at];
    const uint32_t mapB = (2 * Tn + 7) / 8;
    const uint32_t named = (x9 && Tn == 0) ? 1u : 0u;
    if (named) m.secSym[k] = 0x100u | pl[at + 1];                   // at + 5 <= plen: the byte is there
    const uint32_t wAt = at + 1 + named + mapB + 2 * Tn;
    if (wAt + 4 > plen) { m.bad = 1; return at; }
    const uint32_t W = rd32(pl + wAt);
    if ((unsigned long long)wAt + 4 + 4ull * W > plen) { m.bad = 1; return at; }
    m.secTn[k] = Tn; m.secW[k] = W;
    return wAt + 4 + 4 * W;
}

__global__ void __launch_bounds__(128)
k_dec_parse(const uint8_t *__restrict__ file, const uint64_t *__restrict__ recOff, uint32_t nBlocks,
            uint8_t *__restrict__ scratchAll, size_t scratchPerBlock, bool x9)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlocks) return;
    const uint8_t *rec = file + recOff[b];
    const uint32_t plen = rd32(rec);
    const uint8_t *pl = rec + 4;
    DecMeta m;
    m.T = m.Pc = m.G = m.bad = m.nL = 0;
    for (int k = 0; k < 4; k++) { m.secOff[k] = m.secCnt[k] = m.secTn[k] = m.secW[k] = m.secRaw[k] = m.secSym[k] = 0; }
    if (plen < 4) m.bad = 1;
    else {
        const uint32_t T = rd32(pl);
        m.T = T;
        if (T > kBlockBytes) m.bad = 1;
        else {
            uint32_t at = 4;
            const uint32_t mapB = (T + 7) / 8;
            if (mapB > 1) at = parse_section(m, 0, pl, plen, at, mapB, x9);            // mc:2273-2276
            else { m.secRaw[0] = 1; m.secOff[0] = at; m.secCnt[0] = mapB;
                   if (at + mapB > plen) m.bad = 1; at += mapB; }                    // mc:2277-2280
            if (!m.bad) at = parse_section(m, 1, pl, plen, at, T, x9);                 // mc:2283
            if (!m.bad && at + 4 <= plen) {
                m.Pc = rd32(pl + at); at += 4;                                     // mc:2286
                if (m.Pc > DEC_MAX_PC) m.bad = 1;
            } else m.bad = 1;
            if (!m.bad) at = parse_section(m, 2, pl, plen, at, (kPBits * m.Pc) / 8 + 1, x9);   // mc:2290-2292
            if (!m.bad && at + 4 <= plen) {
                m.G = rd32(pl + at); at += 4;                                      // mc:2302
                if (m.G > 105216 / 4) m.bad = 1;
            } else m.bad = 1;
            if (!m.bad) {
                if (m.G) at = parse_section(m, 3, pl, plen, at, m.G * 4, x9);          // mc:2308
                else { m.secCnt[3] = 0; m.secW[3] = 0; m.secTn[3] = 0; m.secRaw[3] = 0; m.secOff[3] = at; }   // App. B8
            }
        }
    }
    *reinterpret_cast<DecMeta *>(scratchAll + (size_t)b * scratchPerBlock + DEC_OFF_META) = m;
}

// =========================================================================================
// 2: the four Huffman sections, one CTA each
// =========================================================================================
struct SecSmem {
    SpecSmem spec;
    uint16_t lut[DEC_LUT];               // len << 8 | symbol for codes of <= 11 bits, else 0
    uint8_t pair[512];
    uint8_t map[64];
};

__global__ void __launch_bounds__(DEC_THREADS)
k_dec_sections(const uint8_t *__restrict__ file, const uint64_t *__restrict__ recOff,
               uint8_t *__restrict__ scratchAll, size_t scratchPerBlock)
{
    __shared__ SecSmem sm;
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    uint8_t *scr = scratchAll + (size_t)b * scratchPerBlock;
    DecMeta *meta = reinterpret_cast<DecMeta *>(scr + DEC_OFF_META);
    if (meta->bad) return;
    const uint8_t *pl = file + recOff[b] + 4;
    const size_t dstOff = k == 0 ? DEC_OFF_BM : k == 1 ? DEC_OFF_CS : k == 2 ? DEC_OFF_PK : DEC_OFF_GB;
    uint8_t *d = scr + dstOff;
    const uint32_t at = meta->secOff[k], cnt = meta->secCnt[k], Tn = meta->secTn[k], W = meta->secW[k];
    if (meta->secRaw[k]) {
        for (uint32_t i = tid; i < cnt; i += DEC_THREADS) d[i] = pl[at + i];
        return;
    }
    // Tn == 0 (single-symbol stream): nothing is decoded, the zero-filled buffer stands (App. B1) -- unless the
    // section names its symbol (FCX9)
    if (meta->secSym[k]) {
        const uint8_t sym = (uint8_t)meta->secSym[k];
        for (uint32_t i = tid; i < cnt; i += DEC_THREADS) d[i] = sym;
        return;
    }
    if (!(cnt && Tn && W)) return;
    const uint32_t mapB = (2 * Tn + 7) / 8;
    for (uint32_t i = tid; i < mapB; i += DEC_THREADS) sm.map[i] = pl[at + 1 + i];
    for (uint32_t i = tid; i < 2 * Tn; i += DEC_THREADS) sm.pair[i] = pl[at + 1 + mapB + i];
    __syncthreads();
    const uint32_t R = Tn + 1, root = Tn - 1;                    // root = last stored node (mc:945-946)
    build_huff_lut(sm.lut, sm.pair, sm.map, Tn);
    __syncthreads();
    HuffCodec cd{sm.lut, sm.pair, sm.map, Tn, R, root};
    LaneBits<true> lb;
    lb.open(pl + at + 1 + mapB + 2 * Tn + 4, W);
    bool bad = false;
    spec_decode(sm.spec, cd, lb, cnt, [d](uint32_t i, uint32_t sym) { d[i] = (uint8_t)sym; }, bad);
    if (bad) meta->bad = 1;                                      // same value from whoever writes it
}

// =========================================================================================
// 3: match lengths, then the byte-parallel token replay
// =========================================================================================
constexpr uint32_t RP_RING = DEC_THREADS <= 256 ? 16384 : 32768;   // power of two >= 2047 (window) + RP_BYTES (one batch) + 258
constexpr uint32_t RP_MASK = RP_RING - 1;
constexpr uint32_t RP_BYTES = DEC_THREADS <= 512 ? 32u * DEC_THREADS : 16384u;     // output bytes per batch at most (batch-relative positions are 16-bit)
constexpr int RP_PER = 8;                           // tokens a thread gathers: one flag byte
constexpr uint32_t RP_TOKENS = DEC_THREADS * RP_PER;      // 2048
constexpr uint32_t RP_NONE = 0xFFFD, RP_STOP = 0xFFFE, RP_BADTOK = 0xFFFF;
static_assert(RP_BYTES + 2047 + 258 <= RP_RING, "window + batch fit the ring");

struct ReplaySmem {
    SpecSmem spec;
    alignas(16) uint8_t ring[RP_RING];   // the most recent output bytes, addressed by position & RP_MASK
    int16_t src[RP_BYTES + 264];         // per output byte of the batch: batch-relative position it copies from (itself: literal)
    uint16_t tokL[RP_TOKENS];            // match length per token (0: literal only), or an RP_* marker
    uint16_t tokP[RP_TOKENS];            // distance
    uint16_t tokPos[RP_TOKENS];          // batch-relative output position of the token
    uint32_t warpA[DEC_WARPS], warpB[DEC_WARPS];
    uint32_t firstStop, firstBad, nTake, takeBytes, takeMatches, bad;
};

__global__ void __launch_bounds__(DEC_THREADS)
k_dec_replay(uint8_t *__restrict__ out, uint32_t *__restrict__ outLen, uint8_t *__restrict__ scratchAll, size_t scratchPerBlock)
{
    extern __shared__ __align__(16) uint8_t rp_smem_raw[];
    ReplaySmem &sm = *reinterpret_cast<ReplaySmem *>(rp_smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *scr = scratchAll + (size_t)b * scratchPerBlock;
    DecMeta *meta = reinterpret_cast<DecMeta *>(scr + DEC_OFF_META);
    if (meta->bad) { if (tid == 0) outLen[b] = DEC_BAD; return; }
    const uint8_t *bm = scr + DEC_OFF_BM, *cs = scr + DEC_OFF_CS, *pk = scr + DEC_OFF_PK;
    uint16_t *lv = reinterpret_cast<uint16_t *>(scr + DEC_OFF_LV);
    uint8_t *o = out + (size_t)b * kBlockBytes;                    // blocks sit back to back: a full block is 1 MiB
    const uint32_t T = meta->T, Pc = meta->Pc, G = meta->G;

    // ---- the l values (golomb_rice_decode, mc:309-358) -----------------------------------------------------
    uint32_t nL = 0;
    if (G && Pc) {
        GolombCodec gc;
        LaneBits<false> gbits;
        gbits.open(scr + DEC_OFF_GB, G);
        bool gbad = false;
        nL = spec_decode(sm.spec, gc, gbits, Pc, [lv](uint32_t i, uint32_t sym) { lv[i] = (uint16_t)sym; }, gbad);
    }
    if (tid == 0) sm.bad = 0;
    __syncthreads();                                                // lv is read by other lanes from here on

    // ---- token replay (my_LZ77_decompress, mc:1716-1735) -----------------------------------------------------
    uint8_t *ring = sm.ring;
    uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
    uint32_t cur = 0, mBase = 0, flushed = 0, tS = 0;               // CTA-uniform
    while (tS < T) {
        // ---- gather up to 2048 tokens: flag bits -> match ranks (popc + scan) -> length, distance ----------------
        const uint32_t tb = tS + (uint32_t)RP_PER * tid;
        const uint32_t nvalid = tb < T ? min((uint32_t)RP_PER, T - tb) : 0u;
        uint32_t flags = 0xFFu;
        if (nvalid) {
            const uint32_t by = tb >> 3, sh = tb & 7;                // the batch may start inside a flag byte
            flags = (((uint32_t)__ldcg(bm + by) | ((uint32_t)__ldcg(bm + by + 1) << 8)) >> sh) & 0xFFu;
        }
        const uint32_t zmask = ~flags & ((1u << nvalid) - 1u);       // match tokens among mine
        if (tid == 0) { sm.firstStop = RP_TOKENS; sm.firstBad = RP_TOKENS; sm.nTake = RP_TOKENS; sm.takeMatches = 0; }
        const bool anyMatch = __syncthreads_or(zmask != 0);          // (also orders the resets above)
        const uint32_t nTok = min(RP_TOKENS, T - tS);
        if (!anyMatch) {
            // ---- literals only: the output is the c stream itself ---------------------------------------------------
            if (cur + nTok > kBlockBytes) { if (tid == 0) sm.bad = 1; break; }
            for (uint32_t j = tid; j < nTok; j += DEC_THREADS) ring[(cur + j) & RP_MASK] = __ldcg(cs + tS + j);
            __syncthreads();
            const uint32_t newCur = cur + nTok, fe = newCur & ~3u;
            for (uint32_t w = (flushed >> 2) + tid; w < (fe >> 2); w += DEC_THREADS)
                o32[w] = reinterpret_cast<const uint32_t *>(ring)[w & (RP_MASK >> 2)];
            flushed = fe; cur = newCur; tS += nTok;
            __syncthreads();                                         // the ring is rewritten by the next batch
            continue;
        }
        // match ranks: exclusive scan of the per-thread match counts
        const uint32_t mine = (uint32_t)__popc(zmask);
        uint32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += v;
        }
        if (lane == 31) sm.warpA[warp] = inc;
        __syncthreads();
        uint32_t rank = mBase + inc - mine;
#pragma unroll
        for (int w = 0; w < DEC_WARPS; w++) if (w < warp) rank += sm.warpA[w];
        uint32_t Lp[RP_PER / 2], len1sum = 0;                        // the eight lengths / markers, two per register
#pragma unroll
        for (int i = 0; i < RP_PER; i++) {
            uint32_t l = RP_NONE, P = 0;
            if ((uint32_t)i < nvalid) {
                l = 0;
                if ((zmask >> i) & 1u) {
                    const uint32_t m = rank + (uint32_t)__popc(zmask & ((1u << i) - 1u));
                    if (m >= Pc) l = RP_STOP;                        // mc:2337-2340: "Fatal Error", stop before this token
                    else if (m >= nL) l = RP_BADTOK;                 // its Golomb code ran past the stream
                    else {
                        l = __ldcg(lv + m);
                        const uint32_t bp = kPBits * m, by = bp >> 3;   // decombine_bits, mc:1315-1338
                        const uint32_t raw = (uint32_t)__ldcg(pk + by) | ((uint32_t)__ldcg(pk + by + 1) << 8) |
                                             ((uint32_t)__ldcg(pk + by + 2) << 16);
                        P = (raw >> (bp & 7)) & 0x7FFu;
                        if (l > (uint32_t)kMaxLen || P == 0) l = RP_BADTOK;
                    }
                }
            }
            if (l == RP_STOP) atomicMin(&sm.firstStop, (uint32_t)RP_PER * tid + i);
            if (l == RP_BADTOK) atomicMin(&sm.firstBad, (uint32_t)RP_PER * tid + i);
            if (i & 1) Lp[i >> 1] |= l << 16; else Lp[i >> 1] = l;
            sm.tokL[RP_PER * tid + i] = (uint16_t)l;
            sm.tokP[RP_PER * tid + i] = (uint16_t)P;
            len1sum += l < RP_NONE ? l + 1 : 0u;
        }
        // output positions: exclusive scan of len + 1 (tokens behind a stop / bad marker never count: cut below)
        uint32_t incB = len1sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incB, d);
            if (lane >= d) incB += v;
        }
        if (lane == 31) sm.warpB[warp] = incB;
        __syncthreads();                                             // warpB, firstStop, firstBad are complete
        uint32_t pos = incB - len1sum;
#pragma unroll
        for (int w = 0; w < DEC_WARPS; w++) if (w < warp) pos += sm.warpB[w];
        // how many tokens this batch takes: up to the first stop / malformed token, and what fits RP_BYTES of output
        const uint32_t firstStop = sm.firstStop, firstBad = sm.firstBad;
        const uint32_t limitTok = min(min(nTok, firstStop), firstBad);
        {
            uint32_t p = pos;
#pragma unroll
            for (int i = 0; i < RP_PER; i++) {
                const uint32_t idx = (uint32_t)RP_PER * tid + i;
                const uint32_t l = (Lp[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
                const uint32_t l1 = l < RP_NONE ? l + 1 : 0u;
                sm.tokPos[idx] = (uint16_t)min(p, 0xFFFFu);
                if (idx < limitTok && p + l1 > RP_BYTES) atomicMin(&sm.nTake, idx);
                p += l1;
            }
        }
        __syncthreads();
        const uint32_t nTake = min(limitTok, sm.nTake);
        if (nTake == 0) {
            // the batch opens with the stop (the block ends here) or with a malformed token
            if (firstBad == 0 && tid == 0) sm.bad = 1;
            break;
        }
        // bytes and matches of the taken tokens; distances are checked against the bytes decoded so far
        {
            uint32_t myMat = 0;
            bool badLane = false;
#pragma unroll
            for (int i = 0; i < RP_PER; i++) {
                const uint32_t idx = (uint32_t)RP_PER * tid + i;
                if (idx < nTake) {
                    const uint32_t l = (Lp[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
                    if ((zmask >> i) & 1u) {
                        myMat++;
                        if (sm.tokP[idx] > cur + sm.tokPos[idx]) badLane = true;     // reaches before the block's first byte
                    }
                    if (idx == nTake - 1) sm.takeBytes = sm.tokPos[idx] + l + 1;
                }
            }
            if (myMat) atomicAdd(&sm.takeMatches, myMat);
            if (badLane) sm.bad = 1;
        }
        __syncthreads();
        const uint32_t B = sm.takeBytes;
        if (sm.bad) break;
        if (cur + B > kBlockBytes) { if (tid == 0) sm.bad = 1; break; }
        // ---- per byte: where it comes from.  Literals go to the ring at once and point at themselves ------------
        for (uint32_t j = tid; j < nTake; j += DEC_THREADS) {
            const uint32_t l = sm.tokL[j], p = sm.tokP[j], at = sm.tokPos[j];
            for (uint32_t k = 0; k < l; k++) sm.src[at + k] = (int16_t)((int)(at + k) - (int)p);     // mc:1730-1732
            sm.src[at + l] = (int16_t)(at + l);
            ring[(cur + at + l) & RP_MASK] = __ldcg(cs + tS + j);                                    // mc:1734
        }
        __syncthreads();
        // ---- pointer doubling: a byte that copies a copy points at that copy's source, until every pointer leaves
        // the batch (negative) or rests on a literal ----------------------------------------------------------------
        for (;;) {
            bool changed = false;
            for (uint32_t q = tid; q < B; q += DEC_THREADS) {
                const int s = sm.src[q];
                if (s >= 0 && s != (int)q) {
                    const int s2 = sm.src[s];
                    if (s2 != s) { sm.src[q] = (int16_t)s2; changed = true; }
                }
            }
            if (!__syncthreads_or(changed)) break;
        }
        // ---- fetch: sources are final (earlier batches, or literals of this one) -----------------------------------
        for (uint32_t q = tid; q < B; q += DEC_THREADS) {
            const int s = sm.src[q];
            if (s != (int)q) ring[(cur + q) & RP_MASK] = ring[(uint32_t)((int)cur + s) & RP_MASK];
        }
        __syncthreads();
        // ---- flush the finished whole words to HBM -------------------------------------------------------------------
        const uint32_t newCur = cur + B, fe = newCur & ~3u;
        for (uint32_t w = (flushed >> 2) + tid; w < (fe >> 2); w += DEC_THREADS)
            o32[w] = reinterpret_cast<const uint32_t *>(ring)[w & (RP_MASK >> 2)];
        flushed = fe; cur = newCur;
        mBase += sm.takeMatches;
        tS += nTake;
        if (nTake < nTok && nTake == firstStop) break;               // the next token is the stop: the block ends here
        if (nTake < nTok && nTake == firstBad) { if (tid == 0) sm.bad = 1; break; }
        __syncthreads();                                             // ring / token arrays are rewritten by the next batch
    }
    __syncthreads();
    const bool bad = sm.bad != 0;
    if (!bad) for (uint32_t i = flushed + tid; i < cur; i += DEC_THREADS) o[i] = ring[i & RP_MASK];
    if (tid == 0) { outLen[b] = bad ? DEC_BAD : cur; meta->nL = nL; }
}

void decode_init()
{
    cudaFuncSetAttribute(k_dec_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ReplaySmem));
}

int launch_decode(const uint8_t *d_file, const uint64_t *d_recOff, uint32_t nBlocks, uint8_t *d_out,
                  uint32_t *d_outLen, uint8_t *d_scratch, size_t scratchPerBlock, cudaStream_t st, bool x9)
{
    // k_dec_replay needs more than 48 KB of dynamic shared memory: opt in on whatever device is current (a process
    // that only ever decodes never creates a compress context, which is where the other kernels are set up)
    decode_init();
    // the reference decodes into zero-initialised buffers (mc:2270,2282,2291,2305)
    cudaMemsetAsync(d_scratch, 0, scratchPerBlock * nBlocks, st);
    k_dec_parse<<<(nBlocks + 127) / 128, 128, 0, st>>>(d_file, d_recOff, nBlocks, d_scratch, scratchPerBlock, x9);
    k_dec_sections<<<dim3(4, nBlocks), DEC_THREADS, 0, st>>>(d_file, d_recOff, d_scratch, scratchPerBlock);
    k_dec_replay<<<nBlocks, DEC_THREADS, sizeof(ReplaySmem), st>>>(d_out, d_outLen, d_scratch, scratchPerBlock);
    return cudaGetLastError() == cudaSuccess ? 3 : -1;          // a kernel that did not launch must not read as "decoded 0 bytes"
}

} // namespace fcx
