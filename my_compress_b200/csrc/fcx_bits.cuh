// fcx_bits.cuh -- bit readers and the speculative parallel decoder of prefix-coded streams, shared by the
// FCX7 block decoder (fcx_decode.cu) and the FCX8 one (fcx_lz78.cu).
#pragma once
#include "fcx_dev.cuh"

namespace fcx {

#ifndef FCX_DEC_THREADS
#define FCX_DEC_THREADS 512
#endif
constexpr int DEC_THREADS = FCX_DEC_THREADS;   // lanes of a decoder CTA (256 or 512)
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int DEC_LUT_BITS = 11;
constexpr int DEC_LUT = 1 << DEC_LUT_BITS;
constexpr uint32_t DEC_SUB_BITS = 512;                            // bits a lane decodes per trip
constexpr uint32_t DEC_FIN = 0xFFFFFFFFu;                         // "the stream ended inside this subsequence"

__device__ __forceinline__ uint32_t rd32(const uint8_t *p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// =========================================================================================
// speculative parallel decode of one prefix-coded stream by one CTA (DEC_THREADS lanes)
// =========================================================================================
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

struct SpecSmem {
    uint32_t endS[DEC_THREADS];          // bit position at which each lane's subsequence ends (DEC_FIN: stream over)
    uint32_t warpTot[DEC_WARPS];
};

// Every lane takes the next DEC_SUB_BITS bits; lane 0 starts on a code boundary, the others guess.  A lane
// decodes until it crosses into the next lane's bits and publishes where it ended; a lane whose predecessor
// ended somewhere else than where it started decodes again from there, until nobody moves (prefix codes
// fall into step after a few symbols, so this is one or two rounds; by induction over the lanes the
// fixed point is the sequential decode whatever the data).  A scan of the symbol counts then gives
// every lane its output offset for the writing pass.  Returns the symbols produced (<= cnt).
template <class Codec, class Bits, class Emit>
__device__ uint32_t spec_decode(SpecSmem &sm, const Codec &cd, Bits lb, uint32_t cnt, Emit emit, bool &bad)
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
        for (int w = 0; w < DEC_WARPS; w++) { const uint32_t v = sm.warpTot[w]; if (w < warp) off += v; total += v; }
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


// 11-bit look-up table of a serialised byte tree (pair / map as stored in a section header, mc:1021-1066): entry =
// len << 8 | symbol for codes of <= 11 bits, else 0.  CTA-wide; the caller puts a barrier behind it.
__device__ __forceinline__ void build_huff_lut(uint16_t *lut, const uint8_t *pair, const uint8_t *map, uint32_t Tn)
{
    const uint32_t R = Tn + 1, root = Tn - 1;                    // root = last stored node (mc:945-946)
    for (uint32_t idx = threadIdx.x; idx < (uint32_t)DEC_LUT; idx += DEC_THREADS) {
        uint32_t node = root, e = 0;
        for (int bit = 0; bit < DEC_LUT_BITS; bit++) {
            const uint32_t c = 2 * node + ((idx >> bit) & 1u);
            const uint32_t ci = pair[c];
            if (!((map[c >> 3] >> (c & 7)) & 1)) { e = ((uint32_t)(bit + 1) << 8) | ci; break; }
            node = ci - (256 - R);
            if (node >= Tn) break;                               // malformed: the exact walk reports it if it is ever reached
        }
        lut[idx] = (uint16_t)e;
    }
}

} // namespace fcx
