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
// A position then walks the chains of the (at most three) sub-regions overlapping its
// window from the far end towards itself: strict '>' on the length gives the reference's
// farthest-wins tie rule, a candidate is only extended when the byte at the current best
// length matches, and a hit of the cap ends the walk (nothing nearer can beat it).
//
// Chain lengths differ wildly between neighbouring positions (a hot Zipf gram has hundreds
// of candidates, its neighbour none); with one position per lane that left ~5 of 32 lanes
// busy.  So when a tile has long chains its positions are first counting-sorted by estimated
// chain length (bucket populations, half-octave bins, heaviest first) and lanes take
// consecutive entries of that order: neighbouring lanes then walk chains of similar length.
#include "fcx_dev.cuh"

namespace fcx {

constexpr int MT_TILE = 4096;
constexpr int MT_SUB = 1024;
constexpr int MT_HALO = 2048;
constexpr int MT_REGION = MT_TILE + MT_HALO;     // 6144 slots
constexpr int MT_NSUB = MT_REGION / MT_SUB;      // 6
constexpr int MT_HBITS = 10;
constexpr int MT_HB = 1 << MT_HBITS;             // buckets per sub-region
constexpr int MT_LOOK = 272;                     // >= 257 + 8
constexpr int MT_THREADS = 256;
constexpr int MT_PER_THREAD = MT_TILE / MT_THREADS;   // 16
constexpr uint32_t MT_NONE = 0xFFFF;
constexpr int MT_BINS = 32;
constexpr uint32_t MT_SORT_THRESHOLD = 12;       // sort only when some bucket holds more than this

struct MatchSmem {
    uint32_t xs[(MT_REGION + MT_LOOK) / 4];
    uint32_t head[MT_NSUB * MT_HB];              // low 16: lowest slot (0xFFFF none), high 16: population
    uint16_t up[MT_REGION];
    uint16_t order[MT_TILE];
    uint8_t bin[MT_TILE];
    uint32_t binCount[MT_BINS];
    uint32_t binStart[MT_BINS];
    uint32_t maxPop;
};

// byte masks: entry m covers bytes 0..m of an 8-byte little-endian window (m >= 7: all eight)
__constant__ unsigned long long c_byteMask[9] = {
    0xFFull, 0xFFFFull, 0xFFFFFFull, 0xFFFFFFFFull, 0xFFFFFFFFFFull, 0xFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFull,
    0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull};

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
    return __funnelshift_r(lds_u32(al), lds_u32(al + 4), (a & 3u) * 8);
}

__device__ __forceinline__ int lcp_capped(uint32_t xsA, int a, int b, int cap)
{
    int k = 0;
    while (k < cap) {
        const uint32_t d = lds_u32_unaligned(xsA + a + k) ^ lds_u32_unaligned(xsA + b + k);
        if (d) {
            const int m = k + ((__ffs(d) - 1) >> 3);
            return m < cap ? m : cap;
        }
        k += 4;
    }
    return cap;
}

__global__ void __launch_bounds__(MT_THREADS)
k_match(const uint8_t *__restrict__ in, const uint32_t *__restrict__ blkLen,
        uint8_t *__restrict__ lenCode, uint16_t *__restrict__ dist)
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
    for (int k = tid; k < MT_NSUB * MT_HB; k += MT_THREADS) sm.head[k] = MT_NONE;
    if (tid < MT_BINS) sm.binCount[tid] = 0;
    if (tid == 0) sm.maxPop = 0;
    __syncthreads();

    // ---- ascending hash chains, one warp per sub-region ---------------------------------
    if (warp < MT_NSUB) {
        const int sr = warp;
        uint32_t *head = sm.head + sr * MT_HB;
        uint32_t myMax = 0;
        for (int ch = MT_SUB / 32 - 1; ch >= 0; ch--) {
            const int slot = sr * MT_SUB + ch * 32 + lane;
            const int pos = regionBase + slot;
            const bool valid = pos >= 0 && pos + 3 <= n;
            const uint32_t h = hash24(ld32u(sm.xs, slot), MT_HBITS);
            const uint32_t key = valid ? h : (0x80000000u | lane);
            const uint32_t grp = __match_any_sync(0xFFFFFFFFu, key);
            const uint32_t higher = (lane == 31) ? 0u : (grp & ~((2u << lane) - 1u));
            uint32_t old = MT_NONE;
            if (valid) old = head[h];
            __syncwarp();
            if (valid) {
                sm.up[slot] = higher ? (uint16_t)(sr * MT_SUB + ch * 32 + (__ffs(higher) - 1)) : (uint16_t)(old & 0xFFFF);
                if ((grp & ((1u << lane) - 1u)) == 0) {
                    const uint32_t pop = (old >> 16) + __popc(grp);
                    head[h] = (uint32_t)slot | (pop << 16);
                    myMax = max(myMax, pop);
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) myMax = max(myMax, __shfl_xor_sync(0xFFFFFFFFu, myMax, d));
        if (lane == 0) atomicMax(&sm.maxPop, myMax);
    }
    __syncthreads();

    const uint32_t xsA = smem_addr(sm.xs), headA = smem_addr(sm.head), upA = smem_addr(sm.up);
    const int minSlot = -regionBase;                       // slot of block position 0
    const int slotEnd = n - regionBase;                    // slot of block position n
    const int tileSlots = min(MT_TILE, slotEnd - MT_HALO); // positions of this tile that exist
    const bool sorted = sm.maxPop > MT_SORT_THRESHOLD;

    // ---- balance: counting sort of the tile's positions by estimated chain length --------------
    if (sorted) {
        for (int k = 0; k < MT_PER_THREAD; k++) {
            const int idx = k * MT_THREADS + tid;
            const int si = MT_HALO + idx;
            uint32_t bn = 0xFF;
            if (idx < tileSlots) {
                const uint32_t h = hash24(lds_u32_unaligned(xsA + si), MT_HBITS);
                const int lo = max(si - kWindow, minSlot);
                uint32_t est = 0;
                for (int sr = lo >> 10; sr <= (si >> 10); sr++) est += lds_u32(headA + 4 * (sr * MT_HB + h)) >> 16;
                bn = 0;                                    // half-octave bins of the estimate
                if (est) {
                    const int msb = 31 - __clz(est);
                    const int half = msb > 0 ? (int)((est >> (msb - 1)) & 1u) : 0;
                    bn = (uint32_t)min(MT_BINS - 1, 1 + 2 * msb + half);
                }
                sm.bin[idx] = (uint8_t)bn;
            }
            const uint32_t grp = __match_any_sync(0xFFFFFFFFu, bn);
            if (bn != 0xFF && (grp & ((1u << lane) - 1u)) == 0) atomicAdd(&sm.binCount[bn], (uint32_t)__popc(grp));
        }
        __syncthreads();
        if (tid == 0) {                                    // heaviest bin first
            uint32_t acc = 0;
            for (int bn = MT_BINS - 1; bn >= 0; bn--) { sm.binStart[bn] = acc; acc += sm.binCount[bn]; }
        }
        __syncthreads();
        for (int k = 0; k < MT_PER_THREAD; k++) {
            const int idx = k * MT_THREADS + tid;
            const uint32_t bn = idx < tileSlots ? (uint32_t)sm.bin[idx] : 0xFFu;
            const uint32_t grp = __match_any_sync(0xFFFFFFFFu, bn);
            const int leader = __ffs(grp) - 1;
            uint32_t base = 0;
            if (bn != 0xFF && lane == leader) base = atomicAdd(&sm.binStart[bn], (uint32_t)__popc(grp));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (bn != 0xFF) sm.order[base + __popc(grp & ((1u << lane) - 1u))] = (uint16_t)(MT_HALO + idx);
        }
        __syncthreads();
    }

    // ---- per-position walk ----------------------------------------------------------------
    // One flat loop over the chains of all overlapping sub-regions.  A visit loads the
    // candidate's first eight bytes and tests them against a register copy of the position's
    // own eight bytes under a mask covering bytes 0..bestL: a candidate that fails cannot beat
    // the current best and costs ~20 straight-line instructions; only a candidate that passes
    // (a strict improvement, a handful per position) gets its exact length computed, and only
    // matches of eight bytes or more enter the extension loop.  Lanes therefore stay in lockstep
    // for almost every visit.
    const size_t outBase = (size_t)b * kBlockBytes;
    uint32_t xsB = xsA, headB = headA, upB = upA;            // opaque copies: keep the bases in registers
    asm volatile("" : "+r"(xsB), "+r"(headB), "+r"(upB));
    const uint32_t orderA = smem_addr(sm.order);
    for (int k = 0; k < MT_PER_THREAD; k++) {
        const int idx = k * MT_THREADS + tid;
        if (idx >= tileSlots) break;
        const int si = sorted ? (int)lds_u16(orderA + 2 * idx) : MT_HALO + idx;
        const int i = regionBase + si;
        const int cap = min(kMaxLen, n - 1 - i);
        int bestL = kMinLen - 1, bestS = 0;
        if (i > 0 && cap >= kMinLen) {
            const uint32_t xi = xsB + si;
            const uint32_t wsi0 = lds_u32_unaligned(xi), wsi1 = lds_u32_unaligned(xi + 4);
            const uint32_t h = hash24(wsi0, MT_HBITS);
            const int lo = max(si - kWindow, minSlot);
            const int srLast = si >> 10;
            const int sr0 = lo >> 10;
            // heads of the (at most three) chains to walk, loaded up front; -1 = no further chain
            const uint32_t hp = headB + 4 * (sr0 * MT_HB + h);
            int s = (int)(lds_u32(hp) & 0xFFFF);
            int next1 = sr0 + 1 <= srLast ? (int)(lds_u32(hp + 4 * MT_HB) & 0xFFFF) : -1;
            int next2 = sr0 + 2 <= srLast ? (int)(lds_u32(hp + 8 * MT_HB) & 0xFFFF) : -1;
            while (s < lo) s = (int)lds_u16(upB + 2 * s);           // entries before the window (MT_NONE = 0xFFFF ends it)
            uint32_t mLo = 0x00FFFFFFu, mHi = 0u;                   // bytes 0..bestL must agree to beat bestL
            for (;;) {
                if (s >= si) {                                      // chain exhausted (MT_NONE >= si as well)
                    if (next1 < 0) break;
                    s = next1; next1 = next2; next2 = -1;
                    continue;
                }
                const int cur = s;
                s = (int)lds_u16(upB + 2 * cur);
                const uint32_t a = xsB + cur, al = a & ~3u, sh = (a & 3u) * 8;
                const uint32_t w0 = lds_u32(al), w1 = lds_u32(al + 4), w2 = lds_u32(al + 8);
                const uint32_t d0 = __funnelshift_r(w0, w1, sh) ^ wsi0;
                const uint32_t d1 = __funnelshift_r(w1, w2, sh) ^ wsi1;
                if ((d0 & mLo) | (d1 & mHi)) continue;              // cannot beat the current best
                // strict improvement (or, once bestL >= 8, a candidate agreeing on all eight bytes)
                const unsigned long long d = ((unsigned long long)d1 << 32) | d0;
                int m = d ? ((__ffsll((long long)d) - 1) >> 3) : 8;
                if (m == 8 && cap > 8) {
                    if (bestL > 8 && lds_u8(a + bestL) != lds_u8(xi + bestL)) continue;
                    m = 8 + lcp_capped(xsB, cur + 8, si + 8, cap - 8);
                }
                m = min(m, cap);
                if (m > bestL) {
                    bestL = m; bestS = cur;
                    if (m == cap) break;                            // nothing nearer can beat the cap
                    const unsigned long long mk = c_byteMask[min(m, 8)];   // now bytes 0..m must agree
                    mLo = (uint32_t)mk; mHi = (uint32_t)(mk >> 32);
                }
            }
        }
        const bool hit = bestL >= kMinLen;
        lenCode[outBase + i] = hit ? (uint8_t)(bestL - 2) : (uint8_t)0;
        dist[outBase + i] = hit ? (uint16_t)(si - bestS) : (uint16_t)0;
    }
}

int match_tiles_per_block() { return (kBlockBytes + MT_TILE - 1) / MT_TILE; }
size_t match_smem_bytes() { return sizeof(MatchSmem); }

// per device: opt in to > 48 KB of dynamic shared memory
void match_init()
{
    cudaFuncSetAttribute(k_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MatchSmem));
}

void launch_match(const uint8_t *in, const uint32_t *blkLen, uint8_t *lenCode, uint16_t *dist,
                  uint32_t nBlocks, uint32_t maxLen, cudaStream_t st)
{
    dim3 grid((maxLen + MT_TILE - 1) / MT_TILE, nBlocks);
    k_match<<<grid, MT_THREADS, sizeof(MatchSmem), st>>>(in, blkLen, lenCode, dist);
}

} // namespace fcx
