// fcx_match.cu -- per-position longest-match kernel (subsystem 1 of the north star).
//
// Replaces longest_match_sunday + Sunday_Search (mc:1446-1514, mc:1407-1443), evaluated at
// EVERY position of a block instead of only at parse cursors.  Rule:
//   cap = min(258, n-i) - 1;  i == 0 or cap < 3  -> literal
//   among s in [max(0,i-2047), i-1]: maximise min(cap, LCP(x[s..], x[i..])) (overlap allowed),
//   ties -> smallest s (farthest);  result kept only when >= 3.
//
// One CTA owns a tile of 4096 positions and stages tile + 2048 bytes of history + 272 bytes
// of look-ahead in shared memory.  The 6144-slot region is cut into six 1024-slot
// sub-regions; warp w threads the positions of sub-region w into ASCENDING per-hash chains
// (hash of the 3-byte gram) with __match_any_sync, walking its chunks from the top so that
// head[] ends up holding the lowest slot and up[] the next higher slot of the same hash;
// head[] also carries the bucket population.
// The six chains of a bucket are then linked into one ascending chain over the whole region
// (the tail of a sub-region's chain points at the head of the next non-empty one, an empty
// head falls through), so a position walks ONE chain from the first sub-region overlapping
// its window towards itself: strict '>' on the length gives the reference's
// farthest-wins tie rule, a candidate is only extended when the byte at the current best
// length matches, and a hit of the cap ends the walk (nothing nearer can beat it).
//
// Chain lengths differ wildly between neighbouring positions (a hot Zipf gram has hundreds
// of candidates, its neighbour none); with one position per lane that left ~5 of 32 lanes
// busy.  So when a tile has long chains its positions are first counting-sorted by estimated
// chain length (bucket populations, half-octave bins, heaviest first) and lanes take
// consecutive entries of that order: neighbouring lanes then walk chains of similar length.
// Long-match tiles are handled lazily.  A 32-position probe measures the mean match length;
// when it is large, almost every position lies inside a match of an earlier position and will
// never be a parse cursor.  The parse can only enter a tile at offsets 0..257, so such a tile
// evaluates exactly those 258 positions plus a speculative chain start every 128 positions, and
// follows next(p) = p + len + 1 from each until it lands on a position that is already
// evaluated (chains merge after a token or two).  The evaluated set then contains every
// possible entry and is closed under next(), so whatever the true entry turns out to be, the
// parse only ever visits evaluated positions; all others keep a literal placeholder.
#include "fcx_dev.cuh"

namespace fcx {

constexpr int MT_TILE = 4096;
#ifndef FCX_MT_SUBBITS
#define FCX_MT_SUBBITS 10           // slots per sub-region = buckets per sub-region = 2^this (10: six sub-regions, 11: three)
#endif
constexpr int MT_SUBBITS = FCX_MT_SUBBITS;
constexpr int MT_SUB = 1 << MT_SUBBITS;
constexpr int MT_HALO = 2048;
constexpr int MT_REGION = MT_TILE + MT_HALO;     // 6144 slots
constexpr int MT_NSUB = MT_REGION / MT_SUB;      // 6
#ifndef FCX_MT_HBITS
#define FCX_MT_HBITS FCX_MT_SUBBITS   // (six sub-regions with 11 bits drop a resident CTA, 9 bits double the collisions: both measured slower)
#endif
#ifndef FCX_MT_SPEC_STRIDE
#define FCX_MT_SPEC_STRIDE 128   // round 1: 32 best of 16..128; with several lanes per position in the thin rounds (round 2) 128 beats 32..256
#endif
constexpr int MT_HBITS = FCX_MT_HBITS;
constexpr int MT_HB = 1 << MT_HBITS;             // buckets per sub-region
constexpr int MT_LOOK = 272;                     // >= 257 + 8 + 7: the 8-byte visit window, then lcp_capped's look-ahead
constexpr int MT_THREADS = 256;
constexpr int MT_PER_THREAD = MT_TILE / MT_THREADS;   // 16
#ifndef FCX_MT_PROBES
#define FCX_MT_PROBES 32
#endif
constexpr int MT_PROBES = FCX_MT_PROBES;
static_assert(MT_PROBES % 32 == 0, "whole warps probe: the sum below is a full-mask shuffle");                         // positions sampled per tile
#ifndef FCX_MT_PROBE_CAP
#define FCX_MT_PROBE_CAP 16
#endif
constexpr int MT_PROBE_CAP = FCX_MT_PROBE_CAP;                      // probed matches are not followed beyond this
#ifndef FCX_MT_LONG_MEAN
#define FCX_MT_LONG_MEAN 12
#endif
constexpr int MT_LONG_MEAN = FCX_MT_LONG_MEAN;                      // mean probed match length that switches a tile to lazy evaluation
#ifndef FCX_MT_LAZY_SPLIT_LOG2
#define FCX_MT_LAZY_SPLIT_LOG2 3
#endif
constexpr int MT_LAZY_SPLIT_LOG2 = FCX_MT_LAZY_SPLIT_LOG2;            // lazy mode: up to 2^this lanes per position in rounds with few positions
constexpr int MT_SPEC_STRIDE = FCX_MT_SPEC_STRIDE;                    // lazy mode: a speculative chain starts every MT_SPEC_STRIDE positions
constexpr uint32_t MT_NONE = 0xFFFF;
constexpr int MT_DENSE_TRIPS = 15;                // after 64 skipped entries the window's first slots are tried directly
constexpr int MT_BINS = 32;
#ifndef FCX_MT_SORT_THRESHOLD
#define FCX_MT_SORT_THRESHOLD 12
#endif
constexpr uint32_t MT_SORT_THRESHOLD = FCX_MT_SORT_THRESHOLD;       // sort only when some bucket holds more than this

struct MatchSmem {
    uint32_t xs[(MT_REGION + MT_LOOK) / 4];
    uint16_t head[MT_NSUB * MT_HB];              // lowest slot of the bucket (0xFFFF none)
    uint16_t up[MT_REGION];
    uint16_t order[MT_TILE];
    uint8_t pop[MT_NSUB * MT_HB / 2];            // population of bucket pairs (h>>1), saturating at 255: work estimate only
    uint32_t binCount[MT_BINS];
    uint32_t binStart[MT_BINS];
    uint32_t maxPop;
    uint32_t probeSum;
    uint16_t tail4[(MT_NSUB - 1) * MT_HB > MT_TILE ? MT_HB : 2];                       // chain tails of sub-region 4 while the chains are linked (0..3: order[])
};
static_assert((MT_NSUB - 1) * MT_HB <= MT_TILE + MT_HB && MT_NSUB * MT_SUB == MT_REGION, "chain tails of all sub-regions but the last fit order[] + tail4[] (the last needs none)");

// ---- explicit shared-memory loads (32-bit shared addresses held in registers) -------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// unaligned little-endian 32-bit read at shared byte address a
__device__ __forceinline__ uint32_t lds_u32_unaligned(uint32_t a)
{
    const uint32_t al = a & ~3u;
    return __funnelshift_r(lds_u32(al), lds_u32(al + 4), a * 8);      // the funnel shift wraps at 32: (a * 8) & 31 = (a & 3) * 8
}

// length of the common prefix of x[a..] and x[b..], at most cap.  Both streams are read as ALIGNED words with
// the previous word carried in a register (one load + one funnel shift per stream per 4 bytes, two steps per
// trip); reads up to 11 bytes past the last compared byte, inside the look-ahead padding of xs.
__device__ __forceinline__ int lcp_capped(uint32_t xsA, int a, int b, int cap)
{
    const uint32_t pa = xsA + (uint32_t)a, pb = xsA + (uint32_t)b;
    uint32_t aa = pa & ~3u, ab = pb & ~3u;
    const uint32_t sa = pa * 8, sb = pb * 8;            // wrapping funnel shifts: only (p & 3) * 8 counts
    uint32_t wa0 = lds_u32(aa), wb0 = lds_u32(ab);
    int k = 0;
    while (k < cap) {
        const uint32_t wa1 = lds_u32(aa + 4), wb1 = lds_u32(ab + 4);
        const uint32_t wa2 = lds_u32(aa + 8), wb2 = lds_u32(ab + 8);
        const uint32_t d0 = __funnelshift_r(wa0, wa1, sa) ^ __funnelshift_r(wb0, wb1, sb);
        const uint32_t d1 = __funnelshift_r(wa1, wa2, sa) ^ __funnelshift_r(wb1, wb2, sb);
        if (d0 | d1) {
            const int m = k + (d0 ? ((__ffs(d0) - 1) >> 3) : 4 + ((__ffs(d1) - 1) >> 3));
            return m < cap ? m : cap;
        }
        wa0 = wa2; wb0 = wb2;
        aa += 8; ab += 8;
        k += 8;
    }
    return cap;
}

// Longest match of the position at slot si (see the kernel comment for the walk).
struct MatchCtx {
    uint32_t xsB, headB, upB, popB;
    int minSlot;
};

// Candidates are the chain entries with slot in [lo, hi): the whole window [max(si - 2047, block start), si) for a
// position evaluated by one lane, a part of it when several lanes share a position (lazy tiles).
// DENSE (lazy tiles only: that is where runs end up): the shortcut for buckets that hold most slots, see below.
template <bool DENSE = false>
__device__ __forceinline__ void match_one(const MatchCtx &c, int si, int cap, int lo, int hi, int &bestLOut, int &bestSOut)
{
    const uint32_t xsB = c.xsB, upB = c.upB;
    int bestL = kMinLen - 1, bestS = 0;
    const uint32_t xi = xsB + si;
    const uint32_t wsi0 = lds_u32_unaligned(xi), wsi1 = lds_u32_unaligned(xi + 4);
    const uint32_t h = hash24(wsi0, MT_HBITS);
    const int sr0 = lo >> MT_SUBBITS;
    uint32_t mLo = 0x00FFFFFFu, mHi = 0u;               // bytes 0..bestL must agree to beat bestL
    // one ascending chain per bucket over the whole region: enter it at the first sub-region the window touches
    int s = (int)lds_u16(c.headB + 2 * (sr0 * MT_HB + h));
    // entries before the window (MT_NONE = 0xFFFF ends the loop).  A bucket that holds most slots (runs of one byte,
    // short periods) would be stepped through entry by entry, hundreds of dependent loads per position: after a few
    // dozen steps look whether the window's first slots are in this bucket themselves -- a slot with this hash value IS a
    // chain entry, and the first of lo, lo + 1, ... that has it is the first entry at or beyond lo.
    if (DENSE) {
        for (int trips = 0; s < lo; trips++) {
            s = (int)lds_u16(upB + 2 * s);              // four entries per trip: the trip counter costs a quarter
            if (s < lo) s = (int)lds_u16(upB + 2 * s);
            if (s < lo) s = (int)lds_u16(upB + 2 * s);
            if (s < lo) s = (int)lds_u16(upB + 2 * s);
            if (trips == MT_DENSE_TRIPS) {
#pragma unroll
                for (int t = 0; t < 4; t++)
                    if (s < lo && lo + t < hi && hash24(lds_u32_unaligned(xsB + lo + t), MT_HBITS) == h) s = lo + t;
            }
        }
    } else
        while (s < lo) s = (int)lds_u16(upB + 2 * s);
    for (;;) {
        if (s >= hi) break;                             // range exhausted (MT_NONE >= hi as well)
        const int cur = s;
        s = (int)lds_u16(upB + 2 * cur);
        const uint32_t a = xsB + cur, al = a & ~3u, sh = a * 8;    // wrapping funnel shift: (a & 3) * 8
        const uint32_t w0 = lds_u32(al), w1 = lds_u32(al + 4), w2 = lds_u32(al + 8);
        const uint32_t c0 = __funnelshift_r(w0, w1, sh), c1 = __funnelshift_r(w1, w2, sh);
        // (candidate ^ position) & mask in one LOP3 per word; the bare differences are only needed on improvement
        uint32_t t0, t1;
        asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(t0) : "r"(c0), "r"(wsi0), "r"(mLo));
        asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(t1) : "r"(c1), "r"(wsi1), "r"(mHi));
        if (t0 | t1) continue;                          // cannot beat the current best
        const uint32_t d0 = c0 ^ wsi0, d1 = c1 ^ wsi1;
        // The candidate agrees on bytes 0..bestL (all eight once bestL >= 7): below eight bytes that IS a strict
        // improvement, no comparison needed (bestL < cap here, or the walk would have ended at the cap).
        const uint32_t dx = d0 ? d0 : d1;               // word holding the first differing byte
        int m;
        if (dx) {
            m = min((d0 ? 0 : 4) + ((__ffs(dx) - 1) >> 3), cap);
        } else {                                        // eight equal bytes
            m = min(8, cap);
            if (cap > 8) {
                if (bestL > 8 && lds_u8(a + bestL) != lds_u8(xi + bestL)) continue;
                m = 8 + lcp_capped(xsB, cur + 8, si + 8, cap - 8);
            }
            if (m <= bestL) continue;
        }
        bestL = m; bestS = cur;
        if (m == cap) break;                            // nothing nearer can beat the cap
        // now bytes 0..m must agree (all eight from m = 7 on): shifts that clamp at 32 build both mask words
        const uint32_t sh8 = 8u * (uint32_t)min(m, 7) + 8u;      // 8 .. 64 mask bits
        uint32_t t;
        asm("shl.b32 %0, %1, %2;" : "=r"(t) : "r"(0xFFFFFFFFu), "r"(sh8));            // 0 from 32 bits on
        mLo = ~t;
        asm("shr.b32 %0, %1, %2;" : "=r"(mHi) : "r"(0xFFFFFFFFu), "r"(64u - sh8));    // 0 up to 32 bits
    }
    bestLOut = bestL; bestSOut = bestS;
}

// eager mode: every position of the tile, one per lane per trip, in work-sorted order
__device__ __forceinline__ void match_tile_eager(MatchSmem &sm, const MatchCtx &ctx, bool sorted, int tileSlots,
                                                 int regionBase, int n, size_t outBase,
                                                 uint8_t *__restrict__ lenCode, uint16_t *__restrict__ dist)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t xsA = ctx.xsB;
    const int minSlot = ctx.minSlot;
    // ---- balance: counting sort of the positions by estimated chain length -----------------------
    if (sorted) {
        // half-octave bin of a position's estimated chain length (sum of bucket populations)
        auto pos_bin = [&](int idx) -> uint32_t {
            const int si = MT_HALO + idx;
            const uint32_t h = hash24(lds_u32_unaligned(xsA + si), MT_HBITS);
            const int lo = max(si - kWindow, minSlot);
            uint32_t est = 0;
            for (int sr = lo >> MT_SUBBITS; sr <= (si >> MT_SUBBITS); sr++) est += lds_u8(ctx.popB + sr * (MT_HB / 2) + (h >> 1));
            if (!est) return 0u;
            const int msb = 31 - __clz(est);
            const int half = msb > 0 ? (int)((est >> (msb - 1)) & 1u) : 0;
            return (uint32_t)min(MT_BINS - 1, 1 + 2 * msb + half);
        };
        // a thread bins the same 16 positions in both passes: keep the bins in registers (8 bits each)
        uint32_t binPack[MT_PER_THREAD / 4];
#pragma unroll
        for (int k = 0; k < MT_PER_THREAD; k++) {
            const int idx = k * MT_THREADS + tid;
            const uint32_t bn = idx < tileSlots ? pos_bin(idx) : 0xFFu;
            if ((k & 3) == 0) binPack[k >> 2] = 0;
            binPack[k >> 2] |= bn << (8 * (k & 3));
            const uint32_t grp = __match_any_sync(0xFFFFFFFFu, bn);
            if (bn != 0xFF && (grp & ((1u << lane) - 1u)) == 0) atomicAdd(&sm.binCount[bn], (uint32_t)__popc(grp));
        }
        __syncthreads();
        if (tid == 0) {                                    // heaviest bin first
            uint32_t acc = 0;
            for (int bn = MT_BINS - 1; bn >= 0; bn--) { sm.binStart[bn] = acc; acc += sm.binCount[bn]; }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < MT_PER_THREAD; k++) {
            const int idx = k * MT_THREADS + tid;
            const uint32_t bn = (binPack[k >> 2] >> (8 * (k & 3))) & 0xFFu;
            const uint32_t grp = __match_any_sync(0xFFFFFFFFu, bn);
            const int leader = __ffs(grp) - 1;
            uint32_t base = 0;
            if (bn != 0xFF && lane == leader) base = atomicAdd(&sm.binStart[bn], (uint32_t)__popc(grp));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (bn != 0xFF) sm.order[base + __popc(grp & ((1u << lane) - 1u))] = (uint16_t)idx;
        }
        __syncthreads();
    }
    const uint32_t orderA = smem_addr(sm.order);
    for (int k = 0; k < MT_PER_THREAD; k++) {
        const int idx = k * MT_THREADS + tid;
        if (idx >= tileSlots) break;
        const int p = sorted ? (int)lds_u16(orderA + 2 * idx) : idx;
        const int si = MT_HALO + p;
        const int i = regionBase + si;
        const int cap = min(kMaxLen, n - 1 - i);
        int bestL = 0, bestS = 0;
        if (i > 0 && cap >= kMinLen) match_one(ctx, si, cap, max(si - kWindow, minSlot), si, bestL, bestS);
        const bool hit = bestL >= kMinLen;
        lenCode[outBase + i] = hit ? (uint8_t)(bestL - 2) : (uint8_t)0;
        dist[outBase + i] = hit ? (uint16_t)(si - bestS) : (uint16_t)0;
    }
}

// lazy mode (long-match tiles): evaluate the 258 possible parse entries and a speculative start
// every MT_SPEC_STRIDE positions, then follow every chain until it lands on an evaluated position
__device__ __forceinline__ void match_tile_lazy(MatchSmem &sm, const MatchCtx &ctx, int tileSlots,
                                                int regionBase, int n, size_t outBase,
                                                uint8_t *__restrict__ lenCode, uint16_t *__restrict__ dist)
{
    const int tid = threadIdx.x;
    // scratch inside order[] (unused in this mode): two work lists and the evaluated bitmap
    uint16_t *listA = sm.order, *listB = sm.order + 1024;
    uint32_t *bits = reinterpret_cast<uint32_t *>(sm.order + 2048);          // 128 words = 4096 bits
    const size_t tileOut = outBase + (size_t)(regionBase + MT_HALO);
    // literal placeholder everywhere; evaluated positions overwrite it after the barrier
    reinterpret_cast<uint4 *>(lenCode + tileOut)[tid] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4 *>(dist + tileOut)[tid] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4 *>(dist + tileOut)[tid + MT_THREADS] = make_uint4(0, 0, 0, 0);
    if (tid < 128) bits[tid] = 0;
    constexpr int kEntries = kParseEntries;                                  // 258
    constexpr int kFirstSpec = (kEntries + MT_SPEC_STRIDE - 1) / MT_SPEC_STRIDE;   // first speculative start beyond the entries
    constexpr int kSpecs = MT_TILE / MT_SPEC_STRIDE - kFirstSpec;
    for (int j = tid; j < kEntries + kSpecs; j += MT_THREADS)
        listA[j] = (uint16_t)(j < kEntries ? j : (kFirstSpec + (j - kEntries)) * MT_SPEC_STRIDE);
    if (tid == 0) { sm.binCount[0] = kEntries + kSpecs; sm.binCount[1] = 0; }
    __syncthreads();
    uint16_t *cur = listA, *nxt = listB;
    for (;;) {
        const uint32_t count = sm.binCount[0];
        if (count == 0) break;
        // A round with few positions gives each of them several lanes, every lane walking its share of the window:
        // the late rounds follow a handful of chains that have not merged yet, one position each, ~50 candidates
        // with long extensions per position, so what the tile waits for there is the latency of ONE evaluation (and a
        // warp instruction costs the same with one lane as with 32).  Full rounds keep one lane per position: the
        // shares cannot prune each other, which costs extensions.  Results are merged with the rule itself:
        // longest, then farthest.
        // (the lanes per position are chosen per trip from what is left of the round: 377 positions run as 256 with
        // one lane each, then 121 with two)
        for (uint32_t base = 0; base < count;) {                                          // CTA-uniform trips (shuffles inside)
            const uint32_t left = count - base;
            const int shift = left > MT_THREADS / 2 ? 0 : left > MT_THREADS / 4 ? 1 : left > MT_THREADS / 8 ? 2 : MT_LAZY_SPLIT_LOG2;
            const int split = 1 << shift;
            const uint32_t g = base + (uint32_t)(tid >> shift);
            base += (uint32_t)(MT_THREADS >> shift);
            const int part = tid & (split - 1);
            const int p = g < count ? (int)cur[g] : tileSlots;
            uint32_t mine = 0;
            if (p < tileSlots && part == 0) {
                const uint32_t bit = 1u << (p & 31);
                mine = (atomicOr(&bits[p >> 5], bit) & bit) ? 0u : 1u;       // else someone else owns this position
            }
            mine = __shfl_sync(0xFFFFFFFFu, mine, (tid & 31) & ~(split - 1));
            const int si = MT_HALO + p;
            const int i = regionBase + si;
            const int cap = min(kMaxLen, n - 1 - i);
            int bestL = 0, bestS = 0;
            if (mine && i > 0 && cap >= kMinLen) {
                const int lo = max(si - kWindow, ctx.minSlot);
                const int q = (si - lo + split - 1) >> shift;
                const int rlo = lo + part * q, rhi = min(si, rlo + q);
                if (rlo < rhi) match_one<true>(ctx, si, cap, rlo, rhi, bestL, bestS);
            }
            uint32_t key = bestL >= kMinLen ? ((uint32_t)bestL << 13) | (0x1FFFu - (uint32_t)bestS) : 0u;
            for (int d = 1; d < split; d <<= 1) key = max(key, __shfl_xor_sync(0xFFFFFFFFu, key, d));
            if (mine && part == 0) {
                const bool hit = key != 0u;
                const int L = (int)(key >> 13), S = 0x1FFF - (int)(key & 0x1FFFu);
                lenCode[outBase + i] = hit ? (uint8_t)(L - 2) : (uint8_t)0;
                dist[outBase + i] = hit ? (uint16_t)(si - S) : (uint16_t)0;
                const int nx = p + (hit ? L : 0) + 1;                        // next parse cursor if p is one
                if (nx < tileSlots && !((bits[nx >> 5] >> (nx & 31)) & 1u)) nxt[atomicAdd(&sm.binCount[1], 1u)] = (uint16_t)nx;
            }
        }
        __syncthreads();
        if (tid == 0) { sm.binCount[0] = sm.binCount[1]; sm.binCount[1] = 0; }
        uint16_t *t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(MT_THREADS)
k_match(const uint8_t *__restrict__ in, const uint32_t *__restrict__ blkLen,
        uint8_t *__restrict__ lenCode, uint16_t *__restrict__ dist, int forceEager)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    MatchSmem &sm = *reinterpret_cast<MatchSmem *>(smem_raw);

    const int b = blockIdx.y;
    const int n = (int)blkLen[b];
    const int tileStart = blockIdx.x * MT_TILE;
    if (tileStart >= n) return;
    const int regionBase = tileStart - MT_HALO;          // block-relative position of slot 0 (may be < 0)
    const uint8_t *x = in + (size_t)b * kBlockBytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- stage bytes: 16-byte vectors where the whole vector is inside [0,n) -------------
    for (int v = tid; v < (MT_REGION + MT_LOOK) / 16; v += MT_THREADS) {
        int pos = regionBase + v * 16;
        uint4 q;
        if (pos >= 0 && pos + 16 <= n) {
            q = *reinterpret_cast<const uint4 *>(x + pos);
        } else {
            uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 16; k++) {
                int p = pos + k;
                if (p >= 0 && p < n) w[k >> 2] |= (uint32_t)x[p] << ((k & 3) * 8);
            }
            q = make_uint4(w[0], w[1], w[2], w[3]);
        }
        reinterpret_cast<uint4 *>(sm.xs)[v] = q;
    }
    for (int k = tid; k < MT_NSUB * MT_HB / 2; k += MT_THREADS) reinterpret_cast<uint32_t *>(sm.head)[k] = 0xFFFFFFFFu;
    for (int k = tid; k < MT_NSUB * MT_HB / 8; k += MT_THREADS) reinterpret_cast<uint32_t *>(sm.pop)[k] = 0u;
    if (tid < MT_BINS) sm.binCount[tid] = 0;
    if (tid == 0) { sm.maxPop = 0; sm.probeSum = 0; }
    __syncthreads();

    // ---- ascending hash chains, one warp per sub-region ---------------------------------
    if (warp < MT_NSUB) {
        const int sr = warp;
        uint16_t *head = sm.head + sr * MT_HB;
        uint8_t *pop = sm.pop + sr * (MT_HB / 2);
        uint16_t *tail = (sr + 1) * MT_HB <= MT_TILE ? sm.order + sr * MT_HB : sm.tail4;   // highest slot of every bucket of this sub-region (not kept for the last one)
        uint32_t myMax = 0;
        for (int ch = MT_SUB / 32 - 1; ch >= 0; ch--) {
            const int slot = sr * MT_SUB + ch * 32 + lane;
            const int pos = regionBase + slot;
            const bool valid = pos >= 0 && pos + 3 <= n;
            const uint32_t h = hash24(ld32u(sm.xs, slot), MT_HBITS);
            // Lanes of this step that share a hash value must be threaded in slot order, which needs the groups
            // (MATCH.ANY) -- a slow instruction: a clock64 trace put a step at ~1700 cycles, and hashing several steps
            // ahead so that their MATCH.ANYs overlap made the kernel slower, i.e. they queue on one unit.  On
            // incompressible input most steps have no two lanes in one bucket: every lane stores its slot in head[h] and
            // reads it back; if everybody reads its own slot there was no collision and each lane is its own group
            // (uniform-256MiB: 3.48 -> 3.24 ms).  head[h] is rewritten properly below either way; `old` is read first.
            uint32_t old = MT_NONE, oldPop = 0;
            if (valid) { old = head[h]; oldPop = pop[h >> 1]; }
            __syncwarp();
            if (valid) head[h] = (uint16_t)slot;
            __syncwarp();
            const bool lost = valid && head[h] != (uint16_t)slot;
            uint32_t grp = 1u << lane;
            if (__any_sync(0xFFFFFFFFu, lost)) grp = __match_any_sync(0xFFFFFFFFu, valid ? h : (0x80000000u | lane));   // (grouping only the colliding lanes: same speed)
            const uint32_t higher = (lane == 31) ? 0u : (grp & ~((2u << lane) - 1u));
            __syncwarp();
            if (valid) {
                sm.up[slot] = higher ? (uint16_t)(sr * MT_SUB + ch * 32 + (__ffs(higher) - 1)) : (uint16_t)old;
                if (!higher && old == MT_NONE && sr < MT_NSUB - 1) tail[h] = (uint16_t)slot;   // chunks come from the top: the bucket's first entry is its tail
                if ((grp & ((1u << lane) - 1u)) == 0) {
                    const uint32_t np = min(255u, oldPop + (uint32_t)__popc(grp));
                    head[h] = (uint16_t)slot;
                    pop[h >> 1] = (uint8_t)np;          // racy against the pair bucket's leader in this step: estimate only
                    myMax = max(myMax, np);
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) myMax = max(myMax, __shfl_xor_sync(0xFFFFFFFFu, myMax, d));
        if (lane == 0) atomicMax(&sm.maxPop, myMax);
    }
    __syncthreads();
    // ---- link the six chains of every bucket into one; an empty (sub-region, bucket) head falls through ----
    for (int h = tid; h < MT_HB; h += MT_THREADS) {
        uint32_t nextHead = sm.head[(MT_NSUB - 1) * MT_HB + h];
        for (int sr = MT_NSUB - 2; sr >= 0; sr--) {
            const uint32_t hd = sm.head[sr * MT_HB + h];
            if (hd != MT_NONE) {
                const uint32_t t = (sr + 1) * MT_HB <= MT_TILE ? sm.order[sr * MT_HB + h] : sm.tail4[h];
                sm.up[t] = (uint16_t)nextHead;
                nextHead = hd;
            } else
                sm.head[sr * MT_HB + h] = (uint16_t)nextHead;
        }
    }
    __syncthreads();

    const uint32_t xsA = smem_addr(sm.xs), headA = smem_addr(sm.head), upA = smem_addr(sm.up);
    const int minSlot = -regionBase;                       // slot of block position 0
    const int slotEnd = n - regionBase;                    // slot of block position n
    const int tileSlots = min(MT_TILE, slotEnd - MT_HALO); // positions of this tile that exist
    const bool sorted = sm.maxPop > MT_SORT_THRESHOLD;

    uint32_t xsB = xsA, headB = headA, upB = upA;            // opaque copies: keep the bases in registers
    asm volatile("" : "+r"(xsB), "+r"(headB), "+r"(upB));
    MatchCtx ctx;
    ctx.xsB = xsB; ctx.headB = headB; ctx.upB = upB; ctx.popB = smem_addr(sm.pop); ctx.minSlot = minSlot;

    // ---- probe: sample 64 positions; a tile whose matches are long is evaluated lazily ---------
    // (only tiles with long chains can be expensive enough to need it: a tile whose every bucket is
    // short is cheap to evaluate eagerly whatever its match lengths are)
    if (sorted && tid < MT_PROBES) {
        const int si = MT_HALO + tid * (MT_TILE / MT_PROBES) + 1;
        int L = 0;
        if (si < slotEnd) {
            const int i = regionBase + si;
            const int cap = min(kMaxLen, n - 1 - i);
            if (i > 0 && cap >= kMinLen) {
                // the probe only has to tell "mean length >= 12" from "short": matches are followed to MT_PROBE_CAP bytes at
                // most (one probe lane walking a 257-byte match kept the other 192 threads at the barrier)
                int bl = 0, bs = 0;
                match_one(ctx, si, min(cap, MT_PROBE_CAP), max(si - kWindow, minSlot), si, bl, bs);
                L = bl >= kMinLen ? bl : 0;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) L += __shfl_xor_sync(0xFFFFFFFFu, L, d);
        if (lane == 0) atomicAdd(&sm.probeSum, (uint32_t)L);
    }
    __syncthreads();
    const size_t outBase = (size_t)b * kBlockBytes;
    if (!forceEager && sm.probeSum >= MT_PROBES * MT_LONG_MEAN)
        match_tile_lazy(sm, ctx, tileSlots, regionBase, n, outBase, lenCode, dist);
    else
        match_tile_eager(sm, ctx, sorted, tileSlots, regionBase, n, outBase, lenCode, dist);
}

int match_tiles_per_block() { return (kBlockBytes + MT_TILE - 1) / MT_TILE; }
size_t match_smem_bytes() { return sizeof(MatchSmem); }

// per device: opt in to > 48 KB of dynamic shared memory
void match_init()
{
    cudaFuncSetAttribute(k_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MatchSmem));
}

// forceEager: evaluate every position (stage probe for tests); the product path passes false
void launch_match(const uint8_t *in, const uint32_t *blkLen, uint8_t *lenCode, uint16_t *dist,
                  uint32_t nBlocks, uint32_t maxLen, bool forceEager, cudaStream_t st)
{
    dim3 grid((maxLen + MT_TILE - 1) / MT_TILE, nBlocks);
    k_match<<<grid, MT_THREADS, sizeof(MatchSmem), st>>>(in, blkLen, lenCode, dist, forceEager ? 1 : 0);
}

} // namespace fcx
