// fcx_api.cu -- the extern "C" boundary (include/fcx.h) and the host-side driver of the
// device pipeline: match -> parse/split -> Huffman plan/hist/tree/layout/encode -> headers.
//
// Host logic mirrors the reference's block loop (mc:4073-4136): 1 MiB blocks, one
// "u32 size | payload" record per block behind a 10-byte header.  Blocks share no state, so
// a multi-GPU call simply gives each device a contiguous block range; the only cross-device
// datum is each shard's byte count, needed to place its records in the output (host prefix sum).
#include "../../include/fcx.h"
#include "fcx_dev.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace fcx {
// fcx_match.cu
void launch_match(const uint8_t *in, const uint32_t *blkLen, uint8_t *lenCode, uint16_t *dist, uint32_t nBlocks,
                  uint32_t maxLen, bool forceEager, cudaStream_t st);
void match_init();
// fcx_parse.cu
int launch_parse(const uint8_t *in, const uint32_t *blkLen, const Ws &ws, uint32_t nBlocks, uint32_t maxLen, cudaStream_t st);
void parse_init();
// fcx_huff.cu
void huff_init();
int launch_huff_stage(const Ws &ws, uint32_t nBlocks, uint8_t *out, uint64_t outCap, uint32_t withHeader,
                      uint32_t hdrTotalBytes, uint32_t hdrBlocks, uint64_t batchStartCursor, int numSM,
                      cudaStream_t st, void (*mark)(void *, int), void *markArg);
// fcx_decode.cu
int launch_decode(const uint8_t *d_file, const uint64_t *d_recOff, uint32_t nBlocks, uint8_t *d_out,
                  uint32_t *d_outLen, uint8_t *d_scratch, size_t scratchPerBlock, cudaStream_t st);
size_t decode_scratch_per_block();
} // namespace fcx

using namespace fcx;

// -----------------------------------------------------------------------------------------
// errors
// -----------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
extern "C" const char *fcx_last_error(void) { return g_err.c_str(); }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(FCX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

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
enum { ST_MATCH = 0, ST_PARSE, ST_PLAN, ST_HIST, ST_TREE, ST_LAYOUT, ST_ENCODE, ST_ENCODE_SLOW, ST_HEADERS, ST_COUNT };
static const char *kStageNames[ST_COUNT] = {"match", "parse", "plan", "hist", "tree", "layout", "encode", "encode_slow", "headers"};

struct fcx_ctx {
    int device = 0;
    int numSM = 148;
    cudaStream_t stream = 0;
    uint32_t maxBlocks = 0;
    uint32_t batchBlocks = 0;          // blocks per device batch (<= maxBlocks); 0 = maxBlocks
    Ws ws{};
    std::vector<void *> allocs;
    uint64_t launches = 0;
    // timing
    bool timing = false;
    std::vector<cudaEvent_t> ev;       // per call: ST_COUNT+1 events per batch
    std::vector<int> evStage;
    float stageMs[ST_COUNT] = {0};
    // last call
    bool pending = false;
    uint64_t lastBytes = 0;
    // decode scratch
    uint8_t *decScratch = nullptr;
    uint64_t *decRecOff = nullptr;
    uint32_t *decOutLen = nullptr;
};

template <typename T>
static cudaError_t dalloc(fcx_ctx *c, T **p, size_t count)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
    if (e == cudaSuccess) { c->allocs.push_back(q); *p = reinterpret_cast<T *>(q); }
    return e;
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
    if (!out || maxBatchBlocks == 0 || maxBatchBlocks > 65535) return fail(FCX_ERR_ARG, "fcx_ctx_create: bad arguments");
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
    A(dalloc(c, &w.parseExit, B * kParseTilesPerBlock * kParseEntries));
    A(dalloc(c, &w.tileEntry, B * kParseTilesPerBlock));
    A(dalloc(c, &w.tileCnt, B * kParseTilesPerBlock));
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
    A(dalloc(c, &w.totals, (size_t)TOT_WORDS));
    A(dalloc(c, &w.counters, (size_t)C_WORDS));
    if (e != cudaSuccess) {
        for (void *p : c->allocs) cudaFree(p);
        delete c;
        return fail(FCX_ERR_CUDA, std::string("workspace allocation: ") + cudaGetErrorString(e));
    }
    cudaMemset(w.totals, 0, TOT_WORDS * sizeof(uint32_t));
    cudaMemset(w.counters, 0, C_WORDS * sizeof(uint64_t));
    match_init();
    parse_init();
    huff_init();
    CK(cudaGetLastError());
    *out = c;
    return FCX_OK;
}

extern "C" void fcx_ctx_destroy(fcx_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    for (void *p : c->allocs) cudaFree(p);
    if (c->decScratch) cudaFree(c->decScratch);
    if (c->decRecOff) cudaFree(c->decRecOff);
    if (c->decOutLen) cudaFree(c->decOutLen);
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
// Optional per-batch hooks of the host-buffer path: batch k starts only after waitBefore[k]
// (its H2D copy), and when it is done the running output cursor is copied to hostCursor[k]
// (pinned) and recordAfter[k] fires, so the host can start that batch's D2H while batch k+1 runs.
struct BatchHooks {
    cudaEvent_t *waitBefore = nullptr;
    cudaEvent_t *recordAfter = nullptr;
    uint64_t *hostCursor = nullptr;
    const uint32_t *sizes = nullptr;        // blocks per batch (uneven pipeline plan); null = the context's uniform step
    uint32_t nSizes = 0;
};

static int run_batches(fcx_ctx *c, const uint8_t *d_in, uint64_t n, uint8_t *d_out, uint64_t outCap, uint32_t flags,
                       uint64_t hdrBytes, uint32_t hdrBlocks, const BatchHooks *hooks = nullptr)
{
    const Ws &w = c->ws;
    cudaStream_t st = c->stream;
    const uint64_t nBlocks = (n + kBlockBytes - 1) / kBlockBytes;
    // edge words of every chunk are merged with atomicOr, so the destination starts out zero
    const uint64_t zeroBytes = std::min<uint64_t>(outCap, fcx_compress_bound(n));
    CK(cudaMemsetAsync(d_out, 0, zeroBytes, st));
    CK(cudaMemsetAsync(w.totals, 0, TOT_WORDS * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(w.counters, 0, C_WORDS * sizeof(uint64_t), st));
    if (nBlocks == 0) {
        if (flags & FCX_WITH_HEADER) {
            // empty input: header only, block_num = 0
            uint8_t h[10] = {'F', 'C', 'X', '7', 0, 0, 0, 0, 0, 0};
            uint32_t tb = (uint32_t)hdrBytes; uint16_t bn = (uint16_t)hdrBlocks;
            memcpy(h + 4, &tb, 4); memcpy(h + 8, &bn, 2);
            if (outCap < 10) return fail(FCX_ERR_CAPACITY, "output buffer too small");
            CK(cudaMemcpyAsync(d_out, h, 10, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
            uint64_t ten = 10;
            CK(cudaMemcpyAsync(w.counters + C_CURSOR, &ten, 8, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
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
        const uint32_t withHeader = (first == 0 && (flags & FCX_WITH_HEADER)) ? 1u : 0u;
        mark_stage(c, -1);
        k_set_blklen<<<(nb + 255) / 256, 256, 0, st>>>(w.blkLen, n, (uint32_t)first, nb);
        launch_match(bin, w.blkLen, w.lenCode, w.dist, nb, maxLen, false, st);
        mark_stage(c, ST_MATCH);
        c->launches += 2 + launch_parse(bin, w.blkLen, w, nb, maxLen, st);
        mark_stage(c, ST_PARSE);
        c->launches += launch_huff_stage(w, nb, d_out, outCap, withHeader, (uint32_t)hdrBytes, hdrBlocks, 0, c->numSM, st,
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
    CK(cudaStreamSynchronize(c->stream));
    if (c->timing) collect_timing(c);
    uint64_t counters[C_WORDS];
    uint32_t totals[TOT_WORDS];
    CK(cudaMemcpy(counters, c->ws.counters, sizeof(counters), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(totals, c->ws.totals, sizeof(totals), cudaMemcpyDeviceToHost));
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
    CK(cudaStreamSynchronize(c->stream));
    uint64_t counters[C_WORDS];
    CK(cudaMemcpy(counters, c->ws.counters, sizeof(counters), cudaMemcpyDeviceToHost));
    s->s_in = counters[C_SIN]; s->s_out = counters[C_SOUT];
    s->tokens = counters[C_TOKENS]; s->matches = counters[C_MATCHES];
    s->container_bytes = counters[C_CURSOR]; s->blocks = counters[C_BLOCKS];
    return FCX_OK;
}

// -----------------------------------------------------------------------------------------
// host-buffer entry points
// -----------------------------------------------------------------------------------------
namespace {
struct DevSlot {
    std::mutex mu;
    fcx_ctx *ctx = nullptr;
    uint32_t ctxBlocks = 0;
    uint8_t *dIn = nullptr, *dOut = nullptr;
    uint64_t inCap = 0, outCap = 0;
    cudaStream_t stream = nullptr;          // kernels
    cudaStream_t h2d = nullptr, d2h = nullptr;
    std::vector<cudaEvent_t> evIn, evOut;   // per sub-batch
    uint64_t *hostCursor = nullptr;         // pinned, per sub-batch
    // decoder scratch (grow-only)
    uint8_t *decScratch = nullptr; uint64_t decScratchCap = 0;
    uint64_t *decRec = nullptr; uint32_t *decLen = nullptr; uint32_t decBlocks = 0;
};
DevSlot g_slots[64];
constexpr uint32_t kMaxPipeBatches = 64;

// grow-only per-device scratch for the host-buffer API
int slot_prepare(DevSlot &s, int dev, uint64_t inBytes, uint64_t outBytes)
{
    CK(cudaSetDevice(dev));
    const uint64_t blocks = std::max<uint64_t>(1, (inBytes + kBlockBytes - 1) / kBlockBytes);
    const uint32_t want = (uint32_t)std::min<uint64_t>(blocks, 1024);
    if (!s.stream) {
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s.h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s.d2h, cudaStreamNonBlocking));
        CK(cudaHostAlloc(reinterpret_cast<void **>(&s.hostCursor), kMaxPipeBatches * sizeof(uint64_t), cudaHostAllocDefault));
        s.evIn.resize(kMaxPipeBatches); s.evOut.resize(kMaxPipeBatches);
        for (uint32_t k = 0; k < kMaxPipeBatches; k++) {
            CK(cudaEventCreateWithFlags(&s.evIn[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s.evOut[k], cudaEventDisableTiming));
        }
    }
    if (!s.ctx || s.ctxBlocks < want) {
        if (s.ctx) fcx_ctx_destroy(s.ctx);
        s.ctx = nullptr;
        int rc = fcx_ctx_create(&s.ctx, dev, want);
        if (rc) return rc;
        s.ctxBlocks = want;
        fcx_ctx_set_stream(s.ctx, s.stream);
    }
    if (s.inCap < inBytes) {
        if (s.dIn) cudaFree(s.dIn);
        s.dIn = nullptr; s.inCap = 0;
        CK(cudaMalloc(&s.dIn, inBytes + 256));
        s.inCap = inBytes;
    }
    if (s.outCap < outBytes) {
        if (s.dOut) cudaFree(s.dOut);
        s.dOut = nullptr; s.outCap = 0;
        CK(cudaMalloc(&s.dOut, outBytes + 256));
        s.outCap = outBytes;
    }
    return FCX_OK;
}
} // namespace

extern "C" int fcx_compress_lz77(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t outCap, uint64_t *outN,
                                 uint32_t deviceMask)
{
    if (!out || !outN || (!in && n)) return fail(FCX_ERR_ARG, "null pointer");      // mc:2122-2123 returns 0 here
    const int nd = fcx_device_count();
    if (nd == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    std::vector<int> devs;
    if (deviceMask == 0) deviceMask = 1;
    for (int g = 0; g < 32 && g < nd; g++) if (deviceMask & (1u << g)) devs.push_back(g);
    if (devs.empty()) return fail(FCX_ERR_ARG, "device_mask selects no visible device");
    const uint64_t nBlocks = (n + kBlockBytes - 1) / kBlockBytes;
    if (outCap < 10) return fail(FCX_ERR_CAPACITY, "output buffer too small");
    const size_t G = std::min<uint64_t>(devs.size(), std::max<uint64_t>(1, nBlocks));
    std::vector<uint64_t> b0(G + 1);
    for (size_t g = 0; g <= G; g++) b0[g] = nBlocks * g / G;                        // contiguous block ranges
    std::vector<uint64_t> bytes(G, 0);
    std::vector<int> rcs(G, FCX_OK);
    std::vector<std::string> errs(G);
    // One shard = one device.  The shard is cut into sub-batches so that the H2D copy of batch k+1
    // and (single-device calls) the D2H copy of batch k-1 overlap the kernels of batch k.
    const bool streamOut = (G == 1);
    auto shard = [&](size_t g) {
        DevSlot &s = g_slots[devs[g]];
        std::lock_guard<std::mutex> lk(s.mu);
        const uint64_t off = b0[g] * (uint64_t)kBlockBytes;
        const uint64_t len = std::min<uint64_t>(n, b0[g + 1] * (uint64_t)kBlockBytes) - std::min<uint64_t>(n, off);
        const uint64_t bound = fcx_compress_bound(len);
        int rc = slot_prepare(s, devs[g], len, bound);
        if (rc) { rcs[g] = rc; errs[g] = g_err; return; }
        const uint64_t blocks = (len + kBlockBytes - 1) / kBlockBytes;
        // Pipeline plan.  Only the first batch's upload and the last batch's download are exposed, so both
        // are small; H2D runs ~3x faster than the kernels, so the second batch can already be five times the
        // first, and the rest goes in a few large batches (every batch pays ~0.15 ms of latency-bound kernels).
#ifndef FCX_PIPE_MIN_BLOCKS
#define FCX_PIPE_MIN_BLOCKS 16
#endif
        std::vector<uint32_t> plan;
        if (blocks > FCX_PIPE_MIN_BLOCKS) {
            const uint32_t B = (uint32_t)blocks, cap = s.ctxBlocks;
            const uint32_t edge = std::min<uint32_t>(16, std::max<uint32_t>(2, B / 16));
            uint32_t rem = B - 2 * edge;
            plan.push_back(edge);
            const uint32_t second = std::min<uint32_t>(rem, 4 * edge + 4);
            if (second) plan.push_back(second);
            rem -= second;
            // large inputs taper at the end as well: the download of a big batch would otherwise outlast the
            // small last batch it is meant to hide behind (incompressible data downloads as much as it uploads)
            const uint32_t taper = (B >= 256 && rem > second) ? second : 0;
            rem -= taper;
            if (rem) {
                const uint32_t lim = std::min<uint32_t>(cap, B >= 256 ? 128u : cap);
                const uint32_t k = std::max<uint32_t>(std::min<uint32_t>(5, std::max<uint32_t>(1, rem / 32)), (rem + lim - 1) / lim);
                for (uint32_t i = 0; i < k; i++) plan.push_back(rem * (i + 1) / k - rem * i / k);
            }
            if (taper) plan.push_back(taper);
            plan.push_back(edge);
        }
        const uint32_t nBatches = (uint32_t)plan.size();
        if (nBatches < 2 || nBatches > kMaxPipeBatches) {                    // small input: nothing to overlap
            s.ctx->batchBlocks = 0;
            if (len && cudaMemcpyAsync(s.dIn, in + off, len, cudaMemcpyHostToDevice, s.stream) != cudaSuccess)
                rc = fail(FCX_ERR_CUDA, "H2D copy failed");
            if (rc == FCX_OK)
                rc = fcx_compress_lz77_device(s.ctx, s.dIn, len, s.dOut, bound, g == 0 ? FCX_WITH_HEADER : 0, n, (uint32_t)nBlocks);
            if (rc == FCX_OK) rc = fcx_ctx_finish(s.ctx, &bytes[g]);
            if (rc == FCX_OK && streamOut) {
                if (bytes[g] > outCap) rc = fail(FCX_ERR_CAPACITY, "output buffer smaller than the compressed result");
                else if (cudaMemcpyAsync(out, s.dOut, bytes[g], cudaMemcpyDeviceToHost, s.stream) != cudaSuccess ||
                         cudaStreamSynchronize(s.stream) != cudaSuccess) rc = fail(FCX_ERR_CUDA, "D2H copy failed");
            }
            rcs[g] = rc;
            if (rc) errs[g] = g_err;
            return;
        }
        cudaError_t e = cudaSuccess;
        uint64_t o = 0;
        for (uint32_t k = 0; k < nBatches && e == cudaSuccess; k++) {
            const uint64_t l = std::min<uint64_t>((uint64_t)plan[k] * kBlockBytes, len - o);
            e = cudaMemcpyAsync(s.dIn + o, in + off + o, l, cudaMemcpyHostToDevice, s.h2d);
            if (e == cudaSuccess) e = cudaEventRecord(s.evIn[k], s.h2d);
            o += l;
        }
        if (e != cudaSuccess) { rcs[g] = fail(FCX_ERR_CUDA, std::string("H2D: ") + cudaGetErrorString(e)); errs[g] = g_err; return; }
        BatchHooks hooks;
        hooks.waitBefore = s.evIn.data(); hooks.recordAfter = s.evOut.data(); hooks.hostCursor = s.hostCursor;
        hooks.sizes = plan.data(); hooks.nSizes = nBatches;
        s.ctx->batchBlocks = 0;
        if (((uintptr_t)s.dIn & 15) || ((uintptr_t)s.dOut & 15)) rc = fail(FCX_ERR_ARG, "device pointers must be 16-byte aligned");
        if (rc == FCX_OK) rc = run_batches(s.ctx, s.dIn, len, s.dOut, bound, g == 0 ? FCX_WITH_HEADER : 0, n, (uint32_t)nBlocks, &hooks);
        s.ctx->batchBlocks = 0;
        if (rc == FCX_OK && streamOut) {
            uint64_t prev = 0;
            for (uint32_t k = 0; k < nBatches && rc == FCX_OK; k++) {
                if (cudaEventSynchronize(s.evOut[k]) != cudaSuccess) { rc = fail(FCX_ERR_CUDA, "batch failed"); break; }
                const uint64_t cur = s.hostCursor[k];
                if (cur > outCap || cur > bound) { rc = fail(FCX_ERR_CAPACITY, "output buffer smaller than the compressed result"); break; }
                if (cur > prev && cudaMemcpyAsync(out + prev, s.dOut + prev, cur - prev, cudaMemcpyDeviceToHost, s.d2h) != cudaSuccess)
                    rc = fail(FCX_ERR_CUDA, "D2H copy failed");
                prev = cur;
            }
        }
        uint64_t got = 0;
        const int frc = fcx_ctx_finish(s.ctx, &got);                         // syncs the kernel stream, reports device-side errors
        if (rc == FCX_OK) rc = frc;
        if (rc == FCX_OK && streamOut && cudaStreamSynchronize(s.d2h) != cudaSuccess) rc = fail(FCX_ERR_CUDA, "D2H copy failed");
        bytes[g] = got;
        rcs[g] = rc;
        if (rc) errs[g] = g_err;
    };
    if (G == 1) shard(0);
    else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++) th.emplace_back(shard, g);
        for (auto &t : th) t.join();
    }
    for (size_t g = 0; g < G; g++) if (rcs[g]) return fail(rcs[g], errs[g]);
    uint64_t total = 0;
    for (size_t g = 0; g < G; g++) total += bytes[g];
    if (total > outCap) return fail(FCX_ERR_CAPACITY, "output buffer smaller than the compressed result");
    if (streamOut) { *outN = total; return FCX_OK; }                     // already copied batch by batch
    // place the shards: only their byte counts cross devices (host prefix sum)
    uint64_t at = 0;
    for (size_t g = 0; g < G; g++) {
        DevSlot &s = g_slots[devs[g]];
        CK(cudaSetDevice(devs[g]));
        CK(cudaMemcpyAsync(out + at, s.dOut, bytes[g], cudaMemcpyDeviceToHost, s.stream));
        at += bytes[g];
    }
    for (size_t g = 0; g < G; g++) { CK(cudaSetDevice(devs[g])); CK(cudaStreamSynchronize(g_slots[devs[g]].stream)); }
    *outN = total;
    return FCX_OK;
}

extern "C" int fcx_compress_block_lz77(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t outCap, uint32_t *outN)
{
    if (!in || !out || !outN) return fail(FCX_ERR_ARG, "null pointer");             // mc:2122-2123
    if (n == 0 || n > kBlockBytes) return fail(FCX_ERR_ARG, "a block holds 1..1048576 bytes (mc:4090)");
    if (fcx_device_count() == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    DevSlot &s = g_slots[0];
    std::lock_guard<std::mutex> lk(s.mu);
    const uint64_t bound = fcx_compress_bound(n);
    int rc = slot_prepare(s, 0, n, bound);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s.dIn, in, n, cudaMemcpyHostToDevice, s.stream));
    rc = fcx_compress_lz77_device(s.ctx, s.dIn, n, s.dOut, bound, 0, 0, 0);
    uint64_t bytes = 0;
    if (rc == FCX_OK) rc = fcx_ctx_finish(s.ctx, &bytes);
    if (rc) return rc;
    if (bytes < 4 || bytes - 4 > outCap) return fail(FCX_ERR_CAPACITY, "output buffer smaller than the payload");
    CK(cudaMemcpy(out, s.dOut + 4, bytes - 4, cudaMemcpyDeviceToHost));             // strip the record's u32 size
    *outN = (uint32_t)(bytes - 4);
    return FCX_OK;
}

// -----------------------------------------------------------------------------------------
// decompress (device kernels in fcx_decode.cu)
// -----------------------------------------------------------------------------------------
static int decode_on_device(int dev, const uint8_t *file, uint64_t nbytes, const std::vector<uint64_t> &recOff,
                            uint8_t *out, uint64_t outCap, uint64_t *outN)
{
    DevSlot &s = g_slots[dev];
    std::lock_guard<std::mutex> lk(s.mu);
    const uint32_t nb = (uint32_t)recOff.size();
    const uint64_t rawCap = (uint64_t)nb * kBlockBytes + 16;       // block b decodes to dOut + b MiB
    int rc = slot_prepare(s, dev, nbytes + 16, rawCap);
    if (rc) return rc;
    fcx_ctx *c = s.ctx;
    const size_t per = decode_scratch_per_block();
    if (s.decScratchCap < per * nb) {
        if (s.decScratch) cudaFree(s.decScratch);
        s.decScratch = nullptr; s.decScratchCap = 0;
        CK(cudaMalloc(&s.decScratch, per * nb + 256));
        s.decScratchCap = per * nb;
    }
    if (s.decBlocks < nb) {
        if (s.decRec) cudaFree(s.decRec);
        if (s.decLen) cudaFree(s.decLen);
        s.decRec = nullptr; s.decLen = nullptr; s.decBlocks = 0;
        CK(cudaMalloc(&s.decRec, sizeof(uint64_t) * (nb + 1)));
        CK(cudaMalloc(&s.decLen, sizeof(uint32_t) * (nb + 1)));
        s.decBlocks = nb;
    }
    CK(cudaMemcpyAsync(s.dIn, file, nbytes, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemsetAsync(s.dIn + nbytes, 0, 16, s.stream));
    CK(cudaMemcpyAsync(s.decRec, recOff.data(), sizeof(uint64_t) * nb, cudaMemcpyHostToDevice, s.stream));
    c->launches += launch_decode(s.dIn, s.decRec, nb, s.dOut, s.decLen, s.decScratch, per, s.stream);
    std::vector<uint32_t> lens(nb);
    CK(cudaMemcpyAsync(lens.data(), s.decLen, sizeof(uint32_t) * nb, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    uint64_t total = 0;
    bool packed = true;                                            // every block but the last is full: one copy
    for (uint32_t b = 0; b < nb; b++) {
        if (lens[b] == 0xFFFFFFFFu) return fail(FCX_ERR_FORMAT, "corrupt block payload");
        if (b + 1 < nb && lens[b] != kBlockBytes) packed = false;
        total += lens[b];
    }
    if (total > outCap) return fail(FCX_ERR_CAPACITY, "output buffer too small");
    if (packed) {
        if (total) CK(cudaMemcpyAsync(out, s.dOut, total, cudaMemcpyDeviceToHost, s.stream));
    } else {
        uint64_t at = 0;
        for (uint32_t b = 0; b < nb; b++) {
            if (lens[b]) CK(cudaMemcpyAsync(out + at, s.dOut + (uint64_t)b * kBlockBytes, lens[b], cudaMemcpyDeviceToHost, s.stream));
            at += lens[b];
        }
    }
    CK(cudaStreamSynchronize(s.stream));
    *outN = total;
    return FCX_OK;
}

extern "C" int fcx_decompress_lz77(const uint8_t *file, uint64_t nbytes, uint8_t *out, uint64_t outCap, uint64_t *outN,
                                   uint32_t deviceMask)
{
    if (!file || !out || !outN) return fail(FCX_ERR_ARG, "null pointer");
    if (nbytes < 10) return fail(FCX_ERR_FORMAT, "shorter than the 10-byte header");               // mc:4139-4144
    if (memcmp(file, "FCX", 3) != 0) return fail(FCX_ERR_FORMAT, "not an FCX container");          // mc:4146-4152
    if (file[3] != '7') return fail(FCX_ERR_UNSUPPORTED, "FCX8 (lz78) containers are not on this path");
    if (fcx_device_count() == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    uint16_t nb16; memcpy(&nb16, file + 8, 2);
    std::vector<uint64_t> rec;
    uint64_t pos = 10;
    for (uint32_t b = 0; b < nb16; b++) {                                                          // mc:4162-4174
        if (pos + 4 > nbytes) return fail(FCX_ERR_FORMAT, "truncated record header");
        uint32_t sz; memcpy(&sz, file + pos, 4);
        if (pos + 4 + sz > nbytes) return fail(FCX_ERR_FORMAT, "truncated record");
        rec.push_back(pos);
        pos += 4 + sz;
    }
    if (rec.empty()) { *outN = 0; return FCX_OK; }
    int dev = 0;
    for (int g = 0; g < 32; g++) if (deviceMask & (1u << g)) { dev = g; break; }
    if (dev >= fcx_device_count()) return fail(FCX_ERR_ARG, "device_mask selects no visible device");
    return decode_on_device(dev, file, nbytes, rec, out, outCap, outN);
}

extern "C" int fcx_decompress_block_lz77(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t outCap,
                                         uint32_t *outN)
{
    if (!payload || !out || !outN) return fail(FCX_ERR_ARG, "null pointer");                      // mc:2257-2258
    if (fcx_device_count() == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    std::vector<uint8_t> rec(4 + (size_t)nbytes);
    memcpy(rec.data(), &nbytes, 4);
    memcpy(rec.data() + 4, payload, nbytes);
    std::vector<uint64_t> off(1, 0);
    uint64_t got = 0;
    int rc = decode_on_device(0, rec.data(), rec.size(), off, out, outCap, &got);
    if (rc) return rc;
    *outN = (uint32_t)got;
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
