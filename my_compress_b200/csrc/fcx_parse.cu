// fcx_parse.cu -- greedy parse + stream split (subsystem 2 of the north star).
//
// Replaces my_LZ77_compress (mc:1675-1714: cursor += l+1), the literal-flag bitmap of
// make_bitMap_table (mc:2073-2094), the c/p/l split (mc:2150-2161), combine_bits at 11 bits
// (mc:1292-1313) and golomb_rice_encode (mc:258-304).  Input is the per-position (len, dist)
// table of k_match; output is the four byte streams the Huffman stage encodes.
//
// The parse is a dependent chain next(i) = i + len(i) + 1 with steps of 1..258.  One CTA
// owns one block and sweeps it in tiles of 4096 positions; inside a tile the chain is
// resolved by a three-level jump table instead of a serial walk:
//   J1[i] = first chain position at or beyond the end of i's 16-position mini-tile
//   J2[i] = first chain position at or beyond the end of i's 256-position mid-tile
// both built by backward sweeps (16 dependent steps each, all threads busy), then 16 serial
// hops give the entry of every mid-tile, 16 more the entry of every mini-tile, and each
// thread walks its own mini-tile (<= 16 tokens).  A CTA-wide scan of (tokens, matches,
// Golomb bits) turns the walk into ranks, so every token scatters its c byte, flag bit,
// 11-bit distance and Golomb code straight to its final bit position -- no serial bit loop.
#include "fcx_dev.cuh"

namespace fcx {

constexpr int PT = 4096;
constexpr int PT_THREADS = 256;
constexpr int PT_LOOK = 272;
constexpr int PT_JPAD = PT + 2 * (PT / 16);

struct ParseSmem {
    uint32_t xs[(PT + PT_LOOK) / 4];
    uint32_t lc[PT / 4];
    uint16_t ds[PT];
    uint16_t J1[PT_JPAD];
    uint16_t J2[PT_JPAD];
    uint16_t midEntry[16];
    uint16_t miniEntry[256];
    uint8_t cbuf[PT + 16];
    uint32_t fbits[PT / 32 + 4];
    uint32_t pbits[368];
    uint32_t gbits[128];
    uint32_t warpTok[8], warpGb[8];
    uint32_t entry;
};

__device__ __forceinline__ int jphys(int i) { return i + 2 * (i >> 4); }
__device__ __forceinline__ int code_len(uint32_t code) { return code ? (int)code + 2 : 0; }

// OR `nbits` (<= 32) low bits of v into a zero-initialised shared bit buffer at bit `pos`
__device__ __forceinline__ void or_bits(uint32_t *buf, uint32_t pos, uint32_t v, int nbits)
{
    if (nbits <= 0) return;
    uint32_t w = pos >> 5, sh = pos & 31;
    atomicOr(&buf[w], v << sh);
    if (sh + nbits > 32) atomicOr(&buf[w + 1], v >> (32 - sh));
}

__global__ void __launch_bounds__(PT_THREADS)
k_parse(const uint8_t *__restrict__ in, const uint32_t *__restrict__ blkLen,
        const uint8_t *__restrict__ lenCode, const uint16_t *__restrict__ dist, Ws ws)
{
    __shared__ ParseSmem sm;
    const int b = blockIdx.x;
    const int n = (int)blkLen[b];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *x = in + (size_t)b * kBlockBytes;
    const uint8_t *lcG = lenCode + (size_t)b * kBlockBytes;
    const uint16_t *dsG = dist + (size_t)b * kBlockBytes;
    uint8_t *fOut = ws.strm[0] + (size_t)b * ws.strmStride[0];
    uint8_t *cOut = ws.strm[1] + (size_t)b * ws.strmStride[1];
    uint8_t *pOut = ws.strm[2] + (size_t)b * ws.strmStride[2];
    uint8_t *gOut = ws.strm[3] + (size_t)b * ws.strmStride[3];
    const uint8_t *lcb = reinterpret_cast<const uint8_t *>(sm.lc);
    const uint8_t *xsb = reinterpret_cast<const uint8_t *>(sm.xs);

    // running totals (identical in every thread)
    uint32_t tokTotal = 0, matTotal = 0, gbitTotal = 0;
    uint32_t fBytes = 0, pBytes = 0, gBytes = 0;   // whole bytes already written
    uint32_t fcar = 0, pcar = 0, gcar = 0;          // carried bits (< 8) sitting at bit 0 of the bit buffers

    for (int k = tid; k < PT / 32 + 4; k += PT_THREADS) sm.fbits[k] = 0;
    for (int k = tid; k < 368; k += PT_THREADS) sm.pbits[k] = 0;
    for (int k = tid; k < 128; k += PT_THREADS) sm.gbits[k] = 0;
    if (tid == 0) sm.entry = 0;
    __syncthreads();

    for (int tb = 0; tb < n; tb += PT) {
        const int limit = min(PT, n - tb);
        // ---- stage the tile --------------------------------------------------------------
        reinterpret_cast<uint4 *>(sm.lc)[tid] = reinterpret_cast<const uint4 *>(lcG + tb)[tid];
        reinterpret_cast<uint4 *>(sm.ds)[tid] = reinterpret_cast<const uint4 *>(dsG + tb)[tid];
        reinterpret_cast<uint4 *>(sm.ds)[tid + 256] = reinterpret_cast<const uint4 *>(dsG + tb)[tid + 256];
        for (int v = tid; v < (PT + PT_LOOK) / 16; v += PT_THREADS) {
            int pos = tb + v * 16;
            uint4 q;
            if (pos + 16 <= n) {
                q = *reinterpret_cast<const uint4 *>(x + pos);
            } else {
                uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                for (int k = 0; k < 16; k++)
                    if (pos + k < n) w[k >> 2] |= (uint32_t)x[pos + k] << ((k & 3) * 8);
                q = make_uint4(w[0], w[1], w[2], w[3]);
            }
            reinterpret_cast<uint4 *>(sm.xs)[v] = q;
        }
        __syncthreads();

        // ---- (a) mini-tile backward sweep -> J1 ------------------------------------------
        {
            const int base = tid * 16;
            uint4 v = reinterpret_cast<const uint4 *>(sm.lc)[tid];
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 15; j >= 0; j--) {
                int i = base + j;
                uint32_t code = (w[j >> 2] >> ((j & 3) * 8)) & 0xFF;
                int len = (i < limit) ? code_len(code) : 0;
                int nx = i + len + 1;
                uint16_t val = (nx >= base + 16) ? (uint16_t)nx : sm.J1[jphys(nx)];
                sm.J1[jphys(i)] = val;
            }
        }
        __syncthreads();
        // ---- (b) mid-tile backward sweep -> J2 (16 lanes per mid-tile, lockstep) -----------
        {
            const int M = tid >> 4, j = tid & 15;
            const int midEnd = (M + 1) * 256;
            for (int m = 15; m >= 0; m--) {
                int i = M * 256 + m * 16 + j;
                int a = sm.J1[jphys(i)];
                uint16_t val = (a >= midEnd) ? (uint16_t)a : sm.J2[jphys(a)];
                sm.J2[jphys(i)] = val;
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- (c) entries of the 16 mid-tiles ------------------------------------------------
        if (tid == 0) {
            int e = (int)sm.entry;
            for (int M = 0; M < 16; M++) {
                sm.midEntry[M] = (uint16_t)e;
                if (e < (M + 1) * 256) e = sm.J2[jphys(e)];
            }
            sm.entry = (uint32_t)(e - PT);     // entry of the next tile (only used when it exists)
        }
        __syncthreads();
        // ---- (d) entries of the 256 mini-tiles -----------------------------------------------
        if (tid < 16) {
            const int M = tid;
            int e = sm.midEntry[M];
            for (int m = 0; m < 16; m++) {
                sm.miniEntry[M * 16 + m] = (uint16_t)e;
                if (e < M * 256 + (m + 1) * 16) e = sm.J1[jphys(e)];
            }
        }
        __syncthreads();
        // ---- (e) walk own mini-tile: count ------------------------------------------------------
        const int hi = min(tid * 16 + 16, limit);
        const int e0 = sm.miniEntry[tid];
        uint32_t myTok = 0, myMat = 0, myGb = 0;
        for (int e = e0; e < hi;) {
            int len = code_len(lcb[e]);
            myTok++;
            if (len) { myMat++; myGb += (uint32_t)(len >> 2) + 3; }
            e += len + 1;
        }
        // CTA exclusive scan of (tok | mat<<16) and gb
        uint32_t packed = myTok | (myMat << 16);
        uint32_t incP = packed, incG = myGb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t tp = __shfl_up_sync(0xFFFFFFFFu, incP, d);
            uint32_t tg = __shfl_up_sync(0xFFFFFFFFu, incG, d);
            if (lane >= d) { incP += tp; incG += tg; }
        }
        if (lane == 31) { sm.warpTok[warp] = incP; sm.warpGb[warp] = incG; }
        __syncthreads();
        uint32_t baseP = 0, baseG = 0, tileP = 0, tileG = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            uint32_t p = sm.warpTok[w], g = sm.warpGb[w];
            if (w < warp) { baseP += p; baseG += g; }
            tileP += p; tileG += g;
        }
        uint32_t exP = baseP + incP - packed, exG = baseG + incG - myGb;
        uint32_t tokIdx = exP & 0xFFFF, matIdx = exP >> 16, gOff = exG;
        const uint32_t tileTok = tileP & 0xFFFF, tileMat = tileP >> 16;

        // ---- emit ------------------------------------------------------------------------------
        for (int e = e0; e < hi;) {
            int len = code_len(lcb[e]);
            sm.cbuf[tokIdx] = xsb[e + len];                      // mc:1489: byte after the match
            if (len == 0) {
                or_bits(sm.fbits, fcar + tokIdx, 1u, 1);         // mc:2087-2089: bit set = literal
            } else {
                or_bits(sm.pbits, pcar + kPBits * matIdx, sm.ds[e], kPBits);   // mc:1301-1309
                uint32_t gp = gcar + gOff;
                int q = len >> 2;                                // mc:267-269: q ones, 0, 2 remainder bits LSB first
                if (q > 32) { or_bits(sm.gbits, gp, 0xFFFFFFFFu, 32); or_bits(sm.gbits, gp + 32, 0xFFFFFFFFu >> (64 - q), q - 32); }
                else if (q > 0) or_bits(sm.gbits, gp, 0xFFFFFFFFu >> (32 - q), q);
                or_bits(sm.gbits, gp + q + 1, (uint32_t)(len & 3), 2);
                matIdx++;
                gOff += (uint32_t)q + 3;
            }
            tokIdx++;
            e += len + 1;
        }
        __syncthreads();

        // ---- flush whole bytes, keep the sub-byte remainder as carry -----------------------------
        for (uint32_t k = tid; k < tileTok; k += PT_THREADS) cOut[tokTotal + k] = sm.cbuf[k];
        const uint32_t fBits = fcar + tileTok, pBitsN = pcar + kPBits * tileMat, gBitsN = gcar + tileG;
        const uint32_t fNb = fBits >> 3, pNb = pBitsN >> 3, gNb = gBitsN >> 3;
        const uint8_t *fb = reinterpret_cast<const uint8_t *>(sm.fbits);
        const uint8_t *pb = reinterpret_cast<const uint8_t *>(sm.pbits);
        const uint8_t *gb = reinterpret_cast<const uint8_t *>(sm.gbits);
        for (uint32_t k = tid; k < fNb; k += PT_THREADS) fOut[fBytes + k] = fb[k];
        for (uint32_t k = tid; k < pNb; k += PT_THREADS) pOut[pBytes + k] = pb[k];
        for (uint32_t k = tid; k < gNb; k += PT_THREADS) gOut[gBytes + k] = gb[k];
        uint32_t carF = 0, carP = 0, carG = 0;
        if (tid == 0) { carF = fb[fNb]; carP = pb[pNb]; carG = gb[gNb]; }
        __syncthreads();
        for (int k = tid; k < PT / 32 + 4; k += PT_THREADS) sm.fbits[k] = 0;
        for (int k = tid; k < 368; k += PT_THREADS) sm.pbits[k] = 0;
        for (int k = tid; k < 128; k += PT_THREADS) sm.gbits[k] = 0;
        __syncthreads();
        if (tid == 0) { sm.fbits[0] = carF; sm.pbits[0] = carP; sm.gbits[0] = carG; }
        tokTotal += tileTok; matTotal += tileMat; gbitTotal += tileG;
        fBytes += fNb; pBytes += pNb; gBytes += gNb;
        fcar = fBits & 7; pcar = pBitsN & 7; gcar = gBitsN & 7;
        __syncthreads();
    }

    // ---- tails ---------------------------------------------------------------------------------
    if (tid == 0) {
        if (fcar) fOut[fBytes++] = (uint8_t)sm.fbits[0];
        pOut[pBytes++] = (uint8_t)sm.pbits[0];                   // (11*Pc)/8 + 1 bytes: partial or spare zero byte (mc:2193)
        const uint32_t G = (gbitTotal + 31) >> 5;                // mc:295-302: last partial word is flushed
        if (gcar) gOut[gBytes++] = (uint8_t)sm.gbits[0];
        while (gBytes < 4 * G) gOut[gBytes++] = 0;
        uint32_t *m = ws.meta + (size_t)b * M_WORDS;
        m[M_T] = tokTotal; m[M_PC] = matTotal; m[M_G] = G; m[M_GBITS] = gbitTotal;
        m[M_S0] = fBytes; m[M_S1] = tokTotal; m[M_S2] = pBytes; m[M_S3] = 4 * G;
    }
}

void launch_parse(const uint8_t *in, const uint32_t *blkLen, const Ws &ws, uint32_t nBlocks, cudaStream_t st)
{
    k_parse<<<nBlocks, PT_THREADS, 0, st>>>(in, blkLen, ws.lenCode, ws.dist, ws);
}

} // namespace fcx
