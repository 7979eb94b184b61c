// fcx_api.cu -- the extern "C" boundary (include/fcx.h) and the host-side driver of the
// device pipeline: match -> parse/split -> Huffman plan/hist/tree/layout/encode -> headers.
//
// Host logic mirrors the reference's block loop (mc:4073-4136): 1 MiB blocks, one
// "u32 size | payload" record per block behind a 10-byte header.  Blocks share no state, so
// a multi-GPU call simply gives each device a contiguous block range; the only cross-device
// datum is each shard's byte count, needed to place its records in the output (host prefix sum).
#include "fcx_internal.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace fcx;

// -----------------------------------------------------------------------------------------
// errors
// -----------------------------------------------------------------------------------------
static thread_local std::string g_err;
namespace fcx {
int fail(int code, const std::string &msg) { g_err = msg; return code; }
const std::string &last_error_string() { return g_err; }
} // namespace fcx
extern "C" const char *fcx_last_error(void) { return g_err.c_str(); }

extern "C" int fcx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" uint64_t fcx_compress_bound(uint64_t n)
{
    const uint64_t nb = (n + kBlockBytes - 1) / kBlockBytes;
    // per block: u32 size + 3 counters + 4 sections x (1+64+510 tree bytes + u32 W + word padding)
    // + c stream <= 8 bit/sym of <= n symbols + bitmap n/8 + p/Golomb never beat that split.
    return 10 + nb * (16 + 4 * 584) + n + n / 8 + 64;
}

// -----------------------------------------------------------------------------------------
// context
// -----------------------------------------------------------------------------------------
static const char *kStageNames[ST_COUNT] = {"match", "parse", "plan", "hist", "tree", "layout", "encode", "encode_slow", "headers"};

constexpr size_t kStatusBytes = C_WORDS * sizeof(uint64_t) + TOT_WORDS * sizeof(uint32_t);
static_assert(TOT_WORDS % 2 == 0, "totals follow the 64-bit counters in one allocation");

// Every workspace array sits between two 256-byte guard bands (fcx_ctx_debug_guards paints / checks them:
// compute-sanitizer is not available on the GPU pool, so a test does the out-of-bounds-write check itself).
constexpr size_t kGuard = 256;
template <typename T>
static cudaError_t dalloc(fcx_ctx *c, T **p, size_t count)
{
    void *q = nullptr;
    const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&q, bytes + 2 * kGuard);
    if (e == cudaSuccess) {
        c->allocs.push_back(q);
        c->guardAt.push_back(static_cast<uint8_t *>(q));
        c->guardAt.push_back(static_cast<uint8_t *>(q) + kGuard + bytes);
        *p = reinterpret_cast<T *>(static_cast<uint8_t *>(q) + kGuard);
    }
    return e;
}

extern "C" int fcx_ctx_debug_guards(fcx_ctx *c, int check)
{
    if (!c) return fail(FCX_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    int bad = 0;
    std::vector<uint8_t> h(kGuard);
    for (uint8_t *g : c->guardAt) {
        if (!check) { CK(cudaMemset(g, 0xA5, kGuard)); continue; }
        CK(cudaMemcpy(h.data(), g, kGuard, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < kGuard; i++) if (h[i] != 0xA5) { bad++; break; }
    }
    return bad;                                             // bands written to since they were painted
}

// Zeroes out[cursor .. cursor + bytes) (clamped to cap), starting at the first 16-byte unit that begins at or
// after the cursor: the bytes before it were zeroed as the previous batch's slack (a bound ends >= 64 bytes
// beyond anything a batch writes).
__global__ void __launch_bounds__(256) k_zero_batch(uint8_t *out, const uint64_t *cursor, uint64_t bytes, uint64_t cap)
{
    const uint64_t cur = *cursor;
    const uint64_t beg = (cur + 15) & ~15ull;
    const uint64_t end = min(cap, cur + bytes);
    if (beg >= end) return;
    const uint64_t nVec = (end - beg) >> 4;
    uint4 *v = reinterpret_cast<uint4 *>(out + beg);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nVec; i += stride) v[i] = make_uint4(0, 0, 0, 0);
    if (blockIdx.x == 0) {
        const uint64_t t0 = beg + (nVec << 4);
        if (t0 + threadIdx.x < end && threadIdx.x < 16) out[t0 + threadIdx.x] = 0;
    }
}

__global__ void k_set_blklen(uint32_t *blkLen, uint64_t n, uint32_t firstBlock, uint32_t nb)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint64_t off = (uint64_t)(firstBlock + b) * kBlockBytes;
    uint64_t rem = n > off ? n - off : 0;
    blkLen[b] = (uint32_t)(rem < kBlockBytes ? rem : kBlockBytes);
}

extern "C" int fcx_ctx_create(fcx_ctx **out, int device, uint32_t maxBatchBlocks)
{
    // a batch holds at most kMaxBatchBlocks blocks: chunk descriptors carry the stream id b*4+k in 12 bits and
    // stream arena offsets are 32-bit (fcx_huff.cu); larger inputs run as several batches on the same context
    if (!out || maxBatchBlocks == 0 || maxBatchBlocks > kMaxBatchBlocks)
        return fail(FCX_ERR_ARG, "fcx_ctx_create: max_batch_blocks must be 1..1024");
    int nd = fcx_device_count();
    if (nd == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    if (device < 0 || device >= nd) return fail(FCX_ERR_ARG, "fcx_ctx_create: no such device");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(FCX_ERR_NO_DEVICE, "kernels are built for sm_100a only");
    fcx_ctx *c = new fcx_ctx();
    c->device = device;
    c->numSM = prop.multiProcessorCount;
    c->maxBlocks = maxBatchBlocks;
    Ws &w = c->ws;
    const size_t B = maxBatchBlocks;
    w.maxBlocks = maxBatchBlocks;
    w.maxChunks = (uint32_t)(B * kMaxChunksPerBlock);
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A(dalloc(c, &w.lenCode, B * kBlockBytes));
    A(dalloc(c, &w.dist, B * kBlockBytes));
    for (int k = 0; k < 4; k++) { w.strmStride[k] = seg_stride(k); A(dalloc(c, &w.strm[k], B * seg_stride(k))); }
    A(dalloc(c, &w.tileExit, B * kParseTilesPerBlock));
    A(dalloc(c, &w.tileState, B * kParseTilesPerBlock));
    A(dalloc(c, &w.blkLen, B));
    A(dalloc(c, &w.meta, B * M_WORDS));
    A(dalloc(c, &w.segHist, B * 4 * 256));
    A(dalloc(c, &w.segTab, B * 4 * 256));
    A(dalloc(c, &w.segHdr, B * 4 * kSegHdrStride));
    A(dalloc(c, &w.segInfo, B * 4 * SI_WORDS));
    A(dalloc(c, &w.segBitBase, B * 4));
    A(dalloc(c, &w.chunkSeg, (size_t)w.maxChunks));
    A(dalloc(c, &w.chunkHist, (size_t)w.maxChunks * 256));
    A(dalloc(c, &w.chunkSlow, (size_t)w.maxChunks));
    A(dalloc(c, &w.chunkBits, (size_t)w.maxChunks));
    A(dalloc(c, &w.chunkBitOff, (size_t)w.maxChunks));
    A(dalloc(c, &w.chunkDesc, (size_t)w.maxChunks));
    A(dalloc(c, &w.segCost, B * 4 + 1));
    A(dalloc(c, &w.encSplit, (size_t)kMaxEncCtas + 1));
    A(dalloc(c, &w.blkOff, B + 1));
    // counters (u64[C_WORDS]) and totals (u32[TOT_WORDS]) share one allocation: fcx_ctx_finish reads both at once
    A(dalloc(c, &w.counters, (size_t)C_WORDS + TOT_WORDS / 2));
    w.totals = reinterpret_cast<uint32_t *>(w.counters + C_WORDS);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&c->hostStatus), kStatusBytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMemset(w.counters, 0, kStatusBytes);
    if (e == cudaSuccess) { match_init(); parse_init(); huff_init(); decode_init(); e = cudaGetLastError(); }
    if (e != cudaSuccess) {
        for (void *p : c->allocs) cudaFree(p);
        if (c->hostStatus) cudaFreeHost(c->hostStatus);
    if (c->decScratch) cudaFree(c->decScratch);
    if (c->decRec) cudaFree(c->decRec);
    if (c->decLen) cudaFree(c->decLen);
        delete c;
        return fail(FCX_ERR_CUDA, std::string("context setup: ") + cudaGetErrorString(e));
    }
    *out = c;
    return FCX_OK;
}

extern "C" void fcx_ctx_destroy(fcx_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    lz78_release(c);
    for (void *p : c->allocs) cudaFree(p);
    if (c->hostStatus) cudaFreeHost(c->hostStatus);
    delete c;
}

extern "C" int fcx_ctx_set_stream(fcx_ctx *c, void *stream)
{
    if (!c) return fail(FCX_ERR_ARG, "null context");
    c->stream = reinterpret_cast<cudaStream_t>(stream);
    return FCX_OK;
}

extern "C" uint64_t fcx_ctx_launch_count(const fcx_ctx *c) { return c ? c->launches : 0; }

extern "C" int fcx_ctx_enable_timing(fcx_ctx *c, int on)
{
    if (!c) return fail(FCX_ERR_ARG, "null context");
    c->timing = on != 0;
    for (int i = 0; i < ST_COUNT; i++) c->stageMs[i] = 0;
    return FCX_OK;
}

static void mark_stage(void *arg, int stageJustFinished)
{
    fcx_ctx *c = reinterpret_cast<fcx_ctx *>(arg);
    if (!c->timing) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    c->ev.push_back(e);
    c->evStage.push_back(stageJustFinished);
}

static void collect_timing(fcx_ctx *c)
{
    // events: a "-1" marker opens each batch; each later event closes the stage it names
    for (size_t i = 1; i < c->ev.size(); i++) {
        if (c->evStage[i] < 0) continue;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[i - 1], c->ev[i]);
        c->stageMs[c->evStage[i]] += ms;
    }
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    c->ev.clear();
    c->evStage.clear();
}

extern "C" int fcx_ctx_stage_times(fcx_ctx *c, const char **names, float *ms, int cap)
{
    if (!c) return fail(FCX_ERR_ARG, "null context");
    for (int i = 0; i < ST_COUNT && i < cap; i++) { if (names) names[i] = kStageNames[i]; if (ms) ms[i] = c->stageMs[i]; }
    return ST_COUNT;
}

// -----------------------------------------------------------------------------------------
// device-resident compress
// -----------------------------------------------------------------------------------------
int fcx::run_batches(fcx_ctx *c, const uint8_t *d_in, uint64_t n, uint8_t *d_out, uint64_t outCap, uint32_t flags,
                     uint64_t hdrBytes, uint32_t hdrBlocks, const BatchHooks *hooks)
{
    const Ws &w = c->ws;
    cudaStream_t st = c->stream;
    const uint64_t nBlocks = (n + kBlockBytes - 1) / kBlockBytes;
    CK(cudaMemsetAsync(w.counters, 0, kStatusBytes, st));           // counters + totals
    if (nBlocks == 0) {
        if (flags & FCX_WITH_HEADER) {
            // empty input: header only, block_num = 0
            uint8_t h[16] = {'F', 'C', 'X', '7', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            uint32_t tb = (uint32_t)hdrBytes; uint16_t bn = (uint16_t)hdrBlocks;
            uint64_t ten = FCX_HEADER_BYTES;
            if (flags & FCX_FORMAT_X9) { h[3] = '9'; memcpy(h + 4, &hdrBytes, 8); memcpy(h + 12, &hdrBlocks, 4); ten = FCX_X9_HEADER_BYTES; }
            else { memcpy(h + 4, &tb, 4); memcpy(h + 8, &bn, 2); }
            if (outCap < ten) return fail(FCX_ERR_CAPACITY, "output buffer too small");
            CK(cudaMemcpyAsync(d_out, h, ten, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(w.counters + C_CURSOR, &ten, 8, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));                              // h / ten are stack variables
        }
        return FCX_OK;
    }
    const uint32_t step = c->batchBlocks ? std::min(c->batchBlocks, c->maxBlocks) : c->maxBlocks;
    uint32_t batch = 0;
    for (uint64_t first = 0, adv = 0; first < nBlocks; first += adv, batch++) {
        const uint32_t want = (hooks && hooks->sizes && batch < hooks->nSizes) ? hooks->sizes[batch] : step;
        const uint32_t nb = (uint32_t)std::min<uint64_t>(std::min(want, c->maxBlocks), nBlocks - first);
        adv = nb;
        if (hooks && hooks->waitBefore) CK(cudaStreamWaitEvent(st, hooks->waitBefore[batch], 0));
        const uint8_t *bin = d_in + first * kBlockBytes;
        const uint64_t rem = n - first * kBlockBytes;
        const uint32_t maxLen = (uint32_t)std::min<uint64_t>(rem, kBlockBytes);
        const uint32_t batchFlags = (flags & FCX_FORMAT_X9) | ((first == 0 && (flags & FCX_WITH_HEADER)) ? FCX_WITH_HEADER : 0u);
        mark_stage(c, -1);
        // edge words of every chunk are merged with atomicOr, so the part of the destination this batch can
        // reach starts out zero (the running cursor lives on the device: the kernel reads it there)
        k_zero_batch<<<c->numSM * 4, 256, 0, st>>>(d_out, w.counters + C_CURSOR, fcx_compress_bound((uint64_t)nb * kBlockBytes), outCap);
        k_set_blklen<<<(nb + 255) / 256, 256, 0, st>>>(w.blkLen, n, (uint32_t)first, nb);
        launch_match(bin, w.blkLen, w.lenCode, w.dist, nb, maxLen, false, st);
        mark_stage(c, ST_MATCH);
        c->launches += 3 + launch_parse(bin, w.blkLen, w, nb, maxLen, st);
        mark_stage(c, ST_PARSE);
        c->launches += launch_huff_stage(w, nb, d_out, outCap, batchFlags, hdrBytes, hdrBlocks, 0, c->numSM, st,
                                         mark_stage, c);
        CK(cudaGetLastError());
        if (hooks && hooks->recordAfter) {
            CK(cudaMemcpyAsync(&hooks->hostCursor[batch], w.counters + C_CURSOR, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(hooks->recordAfter[batch], st));
        }
    }
    return FCX_OK;
}

extern "C" int fcx_compress_lz77_device(fcx_ctx *c, const uint8_t *d_in, uint64_t n, uint8_t *d_out, uint64_t outCap,
                                        uint32_t flags, uint64_t hdrBytes, uint32_t hdrBlocks)
{
    if (!c || !d_out || (!d_in && n)) return fail(FCX_ERR_ARG, "null pointer");
    if (((uintptr_t)d_in & 15) || ((uintptr_t)d_out & 15)) return fail(FCX_ERR_ARG, "device pointers must be 16-byte aligned");
    CK(cudaSetDevice(c->device));
    int rc = run_batches(c, d_in, n, d_out, outCap, flags, hdrBytes, hdrBlocks);
    c->pending = (rc == FCX_OK);
    return rc;
}

extern "C" int fcx_ctx_finish(fcx_ctx *c, uint64_t *outN)
{
    if (!c) return fail(FCX_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    // one pinned read of counters + totals on the context's own stream (no device-wide synchronisation)
    CK(cudaMemcpyAsync(c->hostStatus, c->ws.counters, kStatusBytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->timing) collect_timing(c);
    const uint64_t *counters = c->hostStatus;
    const uint32_t *totals = reinterpret_cast<const uint32_t *>(c->hostStatus + C_WORDS);
    c->pending = false;
    c->lastBytes = counters[C_CURSOR];
    if (outN) *outN = counters[C_CURSOR];
    if (totals[TOT_ERR] & E_CAPACITY) return fail(FCX_ERR_CAPACITY, "output buffer smaller than the compressed result");
    if (totals[TOT_ERR]) return fail(FCX_ERR_CUDA, "device pipeline raised an internal error flag " + std::to_string(totals[TOT_ERR]));
    return FCX_OK;
}

extern "C" int fcx_ctx_stats(fcx_ctx *c, fcx_stats *s)
{
    if (!c || !s) return fail(FCX_ERR_ARG, "null pointer");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->hostStatus, c->ws.counters, kStatusBytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const uint64_t *counters = c->hostStatus;
    s->s_in = counters[C_SIN]; s->s_out = counters[C_SOUT];
    s->tokens = counters[C_TOKENS]; s->matches = counters[C_MATCHES];
    s->container_bytes = counters[C_CURSOR]; s->blocks = counters[C_BLOCKS];
    return FCX_OK;
}

// -----------------------------------------------------------------------------------------
// device-resident decompress
// -----------------------------------------------------------------------------------------
static int decompress_device(fcx_ctx *c, const uint8_t *d_file, const uint64_t *recOff, uint32_t nBlocks, uint8_t *d_out,
                             uint32_t *lens, bool x9);

extern "C" int fcx_decompress_lz77_device(fcx_ctx *c, const uint8_t *d_file, const uint64_t *recOff, uint32_t nBlocks,
                                          uint8_t *d_out, uint32_t *lens)
{
    return decompress_device(c, d_file, recOff, nBlocks, d_out, lens, false);
}

extern "C" int fcx_decompress_lz77_device_x9(fcx_ctx *c, const uint8_t *d_file, const uint64_t *recOff, uint32_t nBlocks,
                                             uint8_t *d_out, uint32_t *lens)
{
    return decompress_device(c, d_file, recOff, nBlocks, d_out, lens, true);
}

static int decompress_device(fcx_ctx *c, const uint8_t *d_file, const uint64_t *recOff, uint32_t nBlocks, uint8_t *d_out,
                             uint32_t *lens, bool x9)
{
    if (!c || !d_file || !recOff || !d_out || !lens) return fail(FCX_ERR_ARG, "null pointer");
    if (nBlocks == 0) return FCX_OK;
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    constexpr uint32_t kBatch = 1024;                     // blocks decoded per launch: bounds the scratch (2.3 MB per block)
    const size_t per = decode_scratch_per_block();
    if (!c->decScratch) {
        CK(cudaMalloc(&c->decScratch, per * kBatch + 256));
        CK(cudaMalloc(&c->decRec, sizeof(uint64_t) * kBatch));
        CK(cudaMalloc(&c->decLen, sizeof(uint32_t) * kBatch));
    }
    for (uint32_t b0 = 0; b0 < nBlocks; b0 += kBatch) {
        const uint32_t nb = std::min(kBatch, nBlocks - b0);
        // pageable source: the copy returns once the offsets are staged, so the caller's array may change afterwards
        CK(cudaMemcpyAsync(c->decRec, recOff + b0, sizeof(uint64_t) * nb, cudaMemcpyHostToDevice, st));
        const int nl = launch_decode(d_file, c->decRec, nb, d_out + (size_t)b0 * kBlockBytes, c->decLen, c->decScratch, per, st, x9);
        if (nl < 0) return fail(FCX_ERR_CUDA, "decode kernels failed to launch");
        c->launches += nl;
        CK(cudaMemcpyAsync(lens + b0, c->decLen, sizeof(uint32_t) * nb, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (uint32_t b = 0; b < nBlocks; b++)
        if (lens[b] == 0xFFFFFFFFu) return fail(FCX_ERR_FORMAT, "corrupt block payload");
    return FCX_OK;
}

// -----------------------------------------------------------------------------------------
// stage probes
// -----------------------------------------------------------------------------------------
__global__ void k_expand_len(const uint8_t *lenCode, uint16_t *len, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) len[i] = lenCode[i] ? (uint16_t)(lenCode[i] + 2) : (uint16_t)0;
}

static int probe_match(fcx_ctx *c, const uint8_t *d_in, uint32_t n, uint16_t *d_len, uint16_t *d_dist, bool eager);

extern "C" int fcx_probe_match(fcx_ctx *c, const uint8_t *d_in, uint32_t n, uint16_t *d_len, uint16_t *d_dist)
{
    return probe_match(c, d_in, n, d_len, d_dist, true);
}

extern "C" int fcx_probe_match_auto(fcx_ctx *c, const uint8_t *d_in, uint32_t n, uint16_t *d_len, uint16_t *d_dist)
{
    return probe_match(c, d_in, n, d_len, d_dist, false);
}

static int probe_match(fcx_ctx *c, const uint8_t *d_in, uint32_t n, uint16_t *d_len, uint16_t *d_dist, bool eager)
{
    if (!c || !d_in || !d_len || !d_dist || n == 0 || n > kBlockBytes) return fail(FCX_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    k_set_blklen<<<1, 256, 0, st>>>(c->ws.blkLen, n, 0, 1);
    launch_match(d_in, c->ws.blkLen, c->ws.lenCode, c->ws.dist, 1, n, eager, st);
    k_expand_len<<<(n + 255) / 256, 256, 0, st>>>(c->ws.lenCode, d_len, n);
    CK(cudaMemcpyAsync(d_dist, c->ws.dist, (size_t)n * 2, cudaMemcpyDeviceToDevice, st));
    c->launches += 3;
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return FCX_OK;
}

extern "C" int fcx_probe_streams(fcx_ctx *c, const uint8_t *d_in, uint32_t n, uint32_t *T, uint32_t *Pc, uint32_t *G,
                                 uint8_t *bitmap, uint8_t *cs, uint8_t *ppack, uint8_t *golomb)
{
    if (!c || !d_in || n == 0 || n > kBlockBytes) return fail(FCX_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    k_set_blklen<<<1, 256, 0, st>>>(c->ws.blkLen, n, 0, 1);
    launch_match(d_in, c->ws.blkLen, c->ws.lenCode, c->ws.dist, 1, n, false, st);
    c->launches += 2 + launch_parse(d_in, c->ws.blkLen, c->ws, 1, n, st);
    uint32_t m[M_WORDS];
    CK(cudaMemcpyAsync(m, c->ws.meta, sizeof(m), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (T) *T = m[M_T];
    if (Pc) *Pc = m[M_PC];
    if (G) *G = m[M_G];
    uint8_t *dst[4] = {bitmap, cs, ppack, golomb};
    for (int k = 0; k < 4; k++)
        if (dst[k] && m[M_S0 + k]) CK(cudaMemcpy(dst[k], c->ws.strm[k], m[M_S0 + k], cudaMemcpyDeviceToHost));
    return FCX_OK;
}

__global__ void k_probe_meta(Ws ws, uint32_t S)
{
    uint32_t *m = ws.meta;
    m[M_T] = S; m[M_PC] = 0; m[M_G] = 0; m[M_GBITS] = 0;
    m[M_S0] = 0; m[M_S1] = S; m[M_S2] = 0; m[M_S3] = 0;
}

extern "C" int fcx_probe_huffman_section(fcx_ctx *c, const uint8_t *d_src, uint32_t S, uint8_t *out, uint32_t outCap,
                                         uint32_t *outN)
{
    if (!c || !d_src || !out || !outN || S > kBlockBytes) return fail(FCX_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const uint64_t cap = fcx_compress_bound(kBlockBytes);
    uint8_t *dOut = nullptr;
    CK(cudaMalloc(&dOut, cap));
    CK(cudaMemsetAsync(dOut, 0, cap, st));
    CK(cudaMemsetAsync(c->ws.totals, 0, TOT_WORDS * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(c->ws.counters, 0, C_WORDS * sizeof(uint64_t), st));
    if (S) CK(cudaMemcpyAsync(c->ws.strm[1], d_src, S, cudaMemcpyDeviceToDevice, st));
    k_probe_meta<<<1, 1, 0, st>>>(c->ws, S);
    c->launches += 1 + launch_huff_stage(c->ws, 1, dOut, cap, 0, 0, 0, 0, c->numSM, st, nullptr, nullptr);
    uint64_t bytes = 0;
    int rc = fcx_ctx_finish(c, &bytes);
    if (rc == FCX_OK) {
        // record = size(4) T(4) sec1 Pc(4) G(4): the probe's other streams are empty
        const uint64_t sec = bytes - 16;
        if (sec > outCap) rc = fail(FCX_ERR_CAPACITY, "output buffer too small");
        else {
            if (sec) cudaMemcpy(out, dOut + 8, sec, cudaMemcpyDeviceToHost);
            *outN = (uint32_t)sec;
        }
    }
    cudaFree(dOut);
    return rc;
}

// -----------------------------------------------------------------------------------------
// corpus generator of config 1 (随机数的生成.cpp:26-42, time() pinned)
// -----------------------------------------------------------------------------------------
extern "C" void fcx_datagen_randm(uint8_t *out, uint32_t m, uint32_t seed)
{
    if (!out || m == 0) return;
    std::vector<int> p(m);
    for (uint32_t i = 0; i < m; i++) p[i] = (int)i;            // gen:31-33
    srand(seed);                                               // the program's first rand() state
    srand(seed + (unsigned)rand());                            // gen:35: srand(time(NULL) + rand())
    uint32_t i = m;
    while (i > 0) {                                            // gen:36-40
        int r = rand() % (int)i;
        --i;
        std::swap(p[i], p[r]);
    }
    for (uint32_t k = 0; k < m; k++) out[k] = (uint8_t)(p[k] & 0xFF);
}
