// fcx_match.cu -- per-position longest-match kernel (subsystem 1 of the north star).
//
// Replaces longest_match_sunday + Sunday_Search (mc:1446-1514, mc:1407-1443), evaluated at
// EVERY position of a block instead of only at parse cursors.  Rule:
//   cap = min(258, n-i) - 1;  i == 0 or cap < 3  -> literal
//   among s in [max(0,i-2047), i-1]: maximise min(cap, LCP(x[s..], x[i..])) (overlap allowed),
//   ties -> smallest s (farthest);  result kept only when >= 3.
//
// One CTA owns a tile of 6144 positions and stages tile + 2048 bytes of history + 272 bytes
// of look-ahead in shared memory.  The 8192-slot region is cut into eight 1024-slot
// sub-regions; warp w threads the positions of sub-region w into ASCENDING per-hash chains
// (hash of the 3-byte gram) with __match_any_sync, walking its chunks from the top so that
// head[] ends up holding the lowest slot and up[] the next higher slot of the same hash.
// A position then walks the chains of the (at most three) sub-regions overlapping its
// window from the far end towards itself: strict '>' on the length gives the reference's
// farthest-wins tie rule, a candidate is only extended when the byte at the current best
// length matches, and a hit of the cap ends the walk (nothing nearer can beat it).
#include "fcx_dev.cuh"

namespace fcx {

constexpr int MT_TILE = 6144;
constexpr int MT_SUB = 1024;
constexpr int MT_HALO = 2048;
constexpr int MT_REGION = MT_TILE + MT_HALO;     // 8192 slots
constexpr int MT_NSUB = MT_REGION / MT_SUB;      // 8
constexpr int MT_HBITS = 11;
constexpr int MT_HB = 1 << MT_HBITS;             // buckets per sub-region
constexpr int MT_LOOK = 272;                     // >= 257 + 8
constexpr int MT_THREADS = 256;
constexpr uint16_t MT_NONE = 0xFFFF;

struct MatchSmem {
    uint32_t xs[(MT_REGION + MT_LOOK) / 4];
    uint16_t head[MT_NSUB * MT_HB];
    uint16_t up[MT_REGION];
};

int match_tiles_per_block() { return (kBlockBytes + MT_TILE - 1) / MT_TILE; }
size_t match_smem_bytes() { return sizeof(MatchSmem); }

__device__ __forceinline__ int lcp_capped(const uint32_t *xs, int a, int b, int cap)
{
    int k = 0;
    while (k < cap) {
        uint32_t d = ld32u(xs, a + k) ^ ld32u(xs, b + k);
        if (d) {
            int m = k + ((__ffs(d) - 1) >> 3);
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
    for (int k = tid; k < MT_NSUB * MT_HB / 2; k += MT_THREADS)
        reinterpret_cast<uint32_t *>(sm.head)[k] = 0xFFFFFFFFu;
    __syncthreads();

    // ---- ascending hash chains, one warp per sub-region ---------------------------------
    {
        const int sr = warp;
        uint16_t *head = sm.head + sr * MT_HB;
        for (int ch = MT_SUB / 32 - 1; ch >= 0; ch--) {
            const int slot = sr * MT_SUB + ch * 32 + lane;
            const int pos = regionBase + slot;
            const bool valid = pos >= 0 && pos + 3 <= n;
            const uint32_t h = hash24(ld32u(sm.xs, slot), MT_HBITS);
            const uint32_t key = valid ? h : (0x80000000u | lane);
            const uint32_t grp = __match_any_sync(0xFFFFFFFFu, key);
            const uint32_t higher = (lane == 31) ? 0u : (grp & ~((2u << lane) - 1u));
            uint16_t old = MT_NONE;
            if (valid) old = head[h];
            __syncwarp();
            if (valid) {
                sm.up[slot] = higher ? (uint16_t)(sr * MT_SUB + ch * 32 + (__ffs(higher) - 1)) : old;
                if ((grp & ((1u << lane) - 1u)) == 0) head[h] = (uint16_t)slot;
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- per-position walk ----------------------------------------------------------------
    const size_t outBase = (size_t)b * kBlockBytes;
    for (int k = 0; k < MT_TILE / MT_THREADS; k++) {
        const int si = MT_HALO + k * MT_THREADS + tid;
        const int i = regionBase + si;
        if (i >= n) break;
        int cap = n - 1 - i;
        if (cap > kMaxLen) cap = kMaxLen;
        int bestL = kMinLen - 1, bestS = 0;
        if (i > 0 && cap >= kMinLen) {
            const uint32_t h = hash24(ld32u(sm.xs, si), MT_HBITS);
            int lo = si - kWindow;
            if (lo < -regionBase) lo = -regionBase;     // block position 0
            const uint8_t *xb = reinterpret_cast<const uint8_t *>(sm.xs);
            bool done = false;
            for (int sr = lo >> 10; sr <= (si >> 10) && !done; sr++) {
                int s = sm.head[sr * MT_HB + h];
                while (s != MT_NONE && s < si) {
                    if (s >= lo && xb[s + bestL] == xb[si + bestL]) {
                        int L = lcp_capped(sm.xs, s, si, cap);
                        if (L > bestL) {
                            bestL = L; bestS = s;
                            if (L == cap) { done = true; break; }
                        }
                    }
                    s = sm.up[s];
                }
            }
        }
        const bool hit = bestL >= kMinLen;
        lenCode[outBase + i] = hit ? (uint8_t)(bestL - 2) : (uint8_t)0;
        dist[outBase + i] = hit ? (uint16_t)(si - bestS) : (uint16_t)0;
    }
}

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
