// fcx_dev.cuh -- constants, workspace layout and small device helpers shared by the kernels.
//
// The constants are part of the FCX7 format (reference: /root/reference/my_compress.cpp,
// "mc:N" = \n-based line N): 1 MiB blocks (mc:113), 2047-byte window (mc:1264), 258-byte
// look-ahead whose last byte is the literal (mc:1265,1459), min match 3 (mc:1267), 11-bit
// distances (mc:1266), Golomb-Rice M=4 (mc:222-224).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fcx {

constexpr uint32_t kBlockBytes = 1u << 20;
constexpr uint32_t kMaxBatchBlocks = 1024;      // blocks per device batch (stream ids b*4+k < 4096, 32-bit arena offsets)
constexpr int kWindow = 2047;
constexpr int kMaxLen = 257;   // min(258, n-i) - 1
constexpr int kMinLen = 3;
constexpr int kPBits = 11;

// ---- stream arena (per block) ---------------------------------------------------------
// seg 0: literal-flag bitmap bytes  ceil(T/8)            <= 131072
// seg 1: c bytes                    T                    <= 1048576
// seg 2: packed 11-bit distances    (11*Pc)/8+1, Pc<=n/4 <= 360449
// seg 3: Golomb words as bytes      4*G, bits <= 0.8*n   <= 104864
constexpr uint32_t kSegStride0 = 131072 + 256;
constexpr uint32_t kSegStride1 = 1048576 + 256;
constexpr uint32_t kSegStride2 = 360704 + 256;
constexpr uint32_t kSegStride3 = 105216 + 256;

// ---- parse tiling -----------------------------------------------------------------------
constexpr int kParseTile = 4096;                                  // positions per parse tile
constexpr int kParseTilesPerBlock = kBlockBytes / kParseTile;     // 256
constexpr int kParseEntries = 258;                                // a chain enters a tile at offset 0..257

// ---- Huffman chunking ------------------------------------------------------------------
constexpr int kChunkSyms = 8192;                  // symbols per encode/histogram chunk
constexpr int kMaxChunksPerBlock = 128 + 17 + 45 + 14;   // c, bitmap, p, golomb (ceil + 1)
constexpr int kSegHdrStride = 576;                // 1 + 64 + 510 rounded up
constexpr int kMaxEncCtas = 4096;                 // upper bound of the k_encode_w grid

// per-block meta words
enum { M_T = 0, M_PC, M_G, M_GBITS, M_S0, M_S1, M_S2, M_S3, M_WORDS };
// per-segment info words
enum { SI_S = 0, SI_R, SI_HDR, SI_W, SI_FIRST_CHUNK, SI_NCHUNKS, SI_RAW, SI_MAXLEN, SI_WORDS };
// totals words
enum { TOT_CHUNKS = 0, TOT_ERR, TOT_SLOW, TOT_WORDS = 8 };
// 64-bit counters
enum { C_CURSOR = 0, C_SIN, C_SOUT, C_TOKENS, C_MATCHES, C_BLOCKS, C_LAST_BYTES, C_WORDS = 8 };

// error bits raised on the device
enum { E_CAPACITY = 1, E_CODELEN = 2, E_INTERNAL = 4 };

struct Ws {
    uint32_t maxBlocks;
    uint32_t maxChunks;
    uint8_t *lenCode;     // [maxBlocks][1 MiB]  0 = literal, else len-2
    uint16_t *dist;       // [maxBlocks][1 MiB]
    uint8_t *strm[4];     // [maxBlocks][stride]
    uint32_t strmStride[4];
    uint32_t *tileExit;   // [maxBlocks][256] bit 31 = known | offset at which the chain enters the NEXT tile (fcx_parse.cu)
    unsigned long long *tileState;   // [maxBlocks][256] look-back status word of each parse tile (fcx_parse.cu)
    uint32_t *blkLen;     // [maxBlocks]
    uint32_t *meta;       // [maxBlocks][M_WORDS]
    uint32_t *segHist;    // [4*maxBlocks][256]
    uint32_t *segTab;     // [4*maxBlocks][256]  code | 1<<len  (0 = symbol absent)
    uint8_t *segHdr;      // [4*maxBlocks][kSegHdrStride]
    uint32_t *segInfo;    // [4*maxBlocks][SI_WORDS]
    uint64_t *segBitBase; // [4*maxBlocks] absolute bit position (in d_out) of code word 0
    uint32_t *chunkSeg;   // [maxChunks]
    uint16_t *chunkHist;  // [maxChunks][256]
    uint8_t *chunkSlow;   // [maxChunks] 1 = codes > 16 bits or too dense for the stage: handled by k_encode_slow
    uint32_t *chunkBits;  // [maxChunks]
    uint32_t *chunkBitOff;// [maxChunks] bit offset inside the segment's code words
    uint4 *chunkDesc;     // [maxChunks] everything k_encode_w needs for one chunk (see CD_* in fcx_huff.cu)
    uint32_t *segCost;    // [4*maxBlocks+1] estimated encode cost per stream, then its exclusive prefix sum
    uint32_t *encSplit;   // [kMaxEncCtas+1] first chunk of every k_encode_w CTA (cost-balanced contiguous ranges)
    uint64_t *blkOff;     // [maxBlocks+1] byte offset (in d_out) of each "u32 size|payload" record
    uint32_t *totals;     // [TOT_WORDS]
    uint64_t *counters;   // [C_WORDS]
};

__host__ __device__ inline uint32_t seg_stride(int k)
{
    return k == 0 ? kSegStride0 : k == 1 ? kSegStride1 : k == 2 ? kSegStride2 : kSegStride3;
}

// ---- device helpers --------------------------------------------------------------------
#ifdef __CUDACC__
// unaligned little-endian 32-bit read from a 4-byte aligned shared/global byte array
__device__ __forceinline__ uint32_t ld32u(const uint32_t *w, int byteOff)
{
    int k = byteOff >> 2;
    return __funnelshift_r(w[k], w[k + 1], (byteOff & 3) * 8);
}

__device__ __forceinline__ uint32_t hash24(uint32_t gram, int bits)
{
    return ((gram & 0xFFFFFFu) * 2654435761u) >> (32 - bits);
}

__device__ __forceinline__ uint32_t hash32(uint32_t gram4, int bits)
{
    return (gram4 * 2654435761u) >> (32 - bits);
}

// byte-granular little-endian stores to arbitrarily aligned global memory
__device__ __forceinline__ void st_u32_bytes(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
#endif

} // namespace fcx
