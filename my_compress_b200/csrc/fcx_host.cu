// fcx_host.cu -- the host-buffer entry points of include/fcx.h: whole files in host memory in, containers
// out, over one or several GPUs.
//
// Host logic mirrors the reference's block loops (compress mc:4073-4136, decompress mc:4138-4202): 1 MiB
// blocks, one "u32 size | payload" record per block behind a 10-byte header.  Blocks share no state, so the
// input is cut into batches of consecutive blocks that are dealt round-robin to the selected devices; there
// is no data-path collective.  The only thing that crosses devices is the byte count of every batch: batch j
// lands in the output right behind batches 0..j-1, so as soon as the sizes of all earlier batches are known
// (a host prefix sum, a few hundred integers per GiB) its bytes travel device -> final position in ONE copy,
// while the devices are already computing their next batches.  Per device three streams run H2D of batch
// k+1, the kernels of batch k and D2H of batch k-1 side by side; only the first upload and the last download
// of a call are exposed, and both are small batches.
#include "fcx_internal.h"

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

using namespace fcx;

namespace {

constexpr uint32_t kMaxPipeBatches = 256;       // batches per device and call
constexpr uint32_t kDecBatchBlocks = 256;       // blocks per decode batch
constexpr uint32_t kDecRing = 3;                // decode batches in flight per device (device output slots)

struct DevSlot {
    std::mutex mu;
    bool streamsUp = false;
    fcx_ctx *ctx = nullptr;
    uint32_t ctxBlocks = 0;
    uint8_t *dIn = nullptr, *dOut = nullptr;
    uint64_t inCap = 0, outCap = 0;
    cudaStream_t stream = nullptr;          // kernels
    cudaStream_t h2d = nullptr, d2h = nullptr;
    std::vector<cudaEvent_t> evIn, evOut, evFree;   // per batch
    uint64_t *hostCursor = nullptr;         // pinned, per batch
    // decoder (grow-only)
    uint8_t *decScratch = nullptr; uint64_t decScratchCap = 0;
    uint64_t *decRec = nullptr; uint32_t *decLen = nullptr; uint64_t decBlocks = 0;
    uint64_t *hostRec = nullptr; uint32_t *hostLen = nullptr; uint64_t hostBlocks = 0;   // pinned mirrors
};
DevSlot g_slots[64];

int slot_streams(DevSlot &s, int dev)
{
    CK(cudaSetDevice(dev));
    if (s.streamsUp) return FCX_OK;
    CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s.h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s.d2h, cudaStreamNonBlocking));
    CK(cudaHostAlloc(reinterpret_cast<void **>(&s.hostCursor), kMaxPipeBatches * sizeof(uint64_t), cudaHostAllocDefault));
    s.evIn.resize(kMaxPipeBatches); s.evOut.resize(kMaxPipeBatches); s.evFree.resize(kMaxPipeBatches);
    for (uint32_t k = 0; k < kMaxPipeBatches; k++) {
        CK(cudaEventCreateWithFlags(&s.evIn[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.evOut[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.evFree[k], cudaEventDisableTiming));
    }
    s.streamsUp = true;
    return FCX_OK;
}

int grow(uint8_t *&p, uint64_t &cap, uint64_t want)
{
    if (cap >= want) return FCX_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    CK(cudaMalloc(&p, want + 256));
    cap = want;
    return FCX_OK;
}

// grow-only per-device scratch of the compress path
int slot_prepare_compress(DevSlot &s, int dev, uint64_t inBytes, uint64_t outBytes, uint32_t maxBatch)
{
    int rc = slot_streams(s, dev);
    if (rc) return rc;
    const uint32_t want = std::max<uint32_t>(1, std::min<uint32_t>(maxBatch, kMaxBatchBlocks));
    if (!s.ctx || s.ctxBlocks < want) {
        if (s.ctx) fcx_ctx_destroy(s.ctx);
        s.ctx = nullptr; s.ctxBlocks = 0;
        rc = fcx_ctx_create(&s.ctx, dev, want);
        if (rc) return rc;
        s.ctxBlocks = want;
        fcx_ctx_set_stream(s.ctx, s.stream);
    }
    if ((rc = grow(s.dIn, s.inCap, inBytes))) return rc;
    return grow(s.dOut, s.outCap, outBytes);
}

// ---- the pipeline plan of one device ---------------------------------------------------------------------
// Only the first batch's upload and the last batch's download are exposed, so both are small; H2D runs ~3x
// faster than the kernels, so the second batch can already be five times the first, and the rest goes in a
// few large batches (every batch pays ~0.15 ms of latency-bound kernels).
#ifndef FCX_PIPE_MIN_BLOCKS
#define FCX_PIPE_MIN_BLOCKS 16
#endif
std::vector<uint32_t> pipeline_plan(uint64_t blocks)
{
    std::vector<uint32_t> plan;
    if (blocks <= FCX_PIPE_MIN_BLOCKS) { if (blocks) plan.push_back((uint32_t)blocks); return plan; }
    const uint64_t B = blocks;
    if (B >= 512) {
        // large inputs: batches double from 8 blocks to 256 (the upload of batch k+1 then fits behind the kernels of
        // batch k from the start: the kernels take ~1.5x as long per MiB as the copy), 256-block batches in the middle,
        // and a tail that shrinks by 2.5x per batch down to 8 so that every download hides behind the next batch's
        // kernels.  Measured on mixed-1GiB against the plan below: 33.96 -> 33.31 ms per call.
        std::vector<uint32_t> up, down;
        uint64_t used = 0;
        for (uint64_t s = 8; s < 256 && used + s <= B / 2; s = s * 5 / 2) { down.push_back((uint32_t)s); used += s; }
        for (uint64_t s = 8; s < 256 && used + s <= B; s *= 2) { up.push_back((uint32_t)s); used += s; }
        // never more batches than the per-device event pool holds: the middle ones grow instead (<= kMaxBatchBlocks each)
        const uint64_t rem = B - used, k = std::min<uint64_t>((rem + 255) / 256, kMaxPipeBatches - 16);
        for (uint32_t s : up) plan.push_back(s);
        for (uint64_t i = 0; i < k; i++) plan.push_back((uint32_t)(rem * (i + 1) / k - rem * i / k));
        for (size_t i = down.size(); i-- > 0;) plan.push_back(down[i]);
        return plan;
    }
    const uint32_t edge = (uint32_t)std::min<uint64_t>(16, std::max<uint64_t>(2, B / 16));
    uint64_t rem = B - 2 * edge;
    plan.push_back(edge);
    const uint32_t second = (uint32_t)std::min<uint64_t>(rem, 4 * edge + 4);
    if (second) plan.push_back(second);
    rem -= second;
    // large inputs taper at the end as well: the download of a big batch would otherwise outlast the small
    // last batch it is meant to hide behind (incompressible data downloads as much as it uploads)
    const uint32_t taper = (B >= 256 && rem > second) ? second : 0;
    rem -= taper;
    if (rem) {
        const uint64_t lim = B >= 256 ? 128u : kMaxBatchBlocks;
        uint64_t k = std::max<uint64_t>(std::min<uint64_t>(5, std::max<uint64_t>(1, rem / 32)), (rem + lim - 1) / lim);
        // never more batches than the per-device event pool holds: grow them instead (<= kMaxBatchBlocks each)
        k = std::min<uint64_t>(k, kMaxPipeBatches - 8);
        for (uint64_t i = 0; i < k; i++) plan.push_back((uint32_t)(rem * (i + 1) / k - rem * i / k));
    }
    if (taper) plan.push_back(taper);
    plan.push_back(edge);
    return plan;
}

// ---- in-order placement of batch outputs ----------------------------------------------------------------------
// publish(j, size) records the size of batch j and returns, once the sizes of all batches before j are known,
// the offset at which batch j starts in the caller's output buffer.
struct Placement {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<uint64_t> size;
    std::vector<char> known;
    std::vector<uint64_t> off;          // off[j] valid once j <= prefix
    size_t prefix = 0;                  // sizes of batches [0, prefix) are known
    bool failed = false;
    explicit Placement(size_t n, uint64_t base) : size(n, 0), known(n, 0), off(n + 1, 0) { off[0] = base; }
    // false: another shard failed, nothing will be placed
    bool publish(size_t j, uint64_t bytes, uint64_t &at)
    {
        std::unique_lock<std::mutex> lk(mu);
        size[j] = bytes; known[j] = 1;
        while (prefix < known.size() && known[prefix]) { off[prefix + 1] = off[prefix] + size[prefix]; prefix++; }
        cv.notify_all();
        cv.wait(lk, [&] { return failed || prefix >= j; });
        if (failed) return false;
        at = off[j];
        return true;
    }
    void fail_all()
    {
        std::lock_guard<std::mutex> lk(mu);
        failed = true;
        cv.notify_all();
    }
    uint64_t total() const { return off[known.size()]; }
};

struct GBatch {
    uint64_t firstBlock;
    uint32_t nBlocks;
    uint32_t dev;       // index into the call's device list
};

std::vector<int> devices_of(uint32_t mask, int nd)
{
    std::vector<int> devs;
    if (mask == 0) mask = 1;
    for (int g = 0; g < 32 && g < nd; g++) if (mask & (1u << g)) devs.push_back(g);
    return devs;
}

// compress `n` bytes into records (+ the file header when withHeader) over the devices of `mask`
// fmt: 0 = FCX7 (the reference's bytes), FCX_FORMAT_X9 = the opt-in format (include/fcx.h)
int compress_host(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t outCap, uint64_t *outN, uint32_t mask,
                  bool withHeader, uint64_t hdrBytes, uint64_t hdrBlocks, uint32_t fmt = 0)
{
    if (!out || !outN || (!in && n)) return fail(FCX_ERR_ARG, "null pointer");      // mc:2122-2123 returns 0 here
    const int nd = fcx_device_count();
    if (nd == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    const std::vector<int> devs = devices_of(mask, nd);
    if (devs.empty()) return fail(FCX_ERR_ARG, "device_mask selects no visible device");
    const uint64_t nBlocks = (n + kBlockBytes - 1) / kBlockBytes;
    if (withHeader && outCap < (fmt ? FCX_X9_HEADER_BYTES : FCX_HEADER_BYTES)) return fail(FCX_ERR_CAPACITY, "output buffer too small");
    const size_t G = (size_t)std::min<uint64_t>(devs.size(), std::max<uint64_t>(1, nBlocks));

    // ---- deal the batches: round k gives every device plan[k] consecutive blocks -----------------------------
    const std::vector<uint32_t> plan = pipeline_plan((nBlocks + G - 1) / G);
    std::vector<GBatch> batches;
    std::vector<std::vector<size_t>> mine(G);
    {
        uint64_t at = 0;
        for (size_t k = 0; k < plan.size() && at < nBlocks; k++)
            for (size_t g = 0; g < G && at < nBlocks; g++) {
                const uint32_t nb = (uint32_t)std::min<uint64_t>(plan[k], nBlocks - at);
                mine[g].push_back(batches.size());
                batches.push_back({at, nb, (uint32_t)g});
                at += nb;
            }
    }
    for (uint32_t sz : plan)
        if (sz > kMaxBatchBlocks) return fail(FCX_ERR_UNSUPPORTED, "input too large for one call: split it (fcx_compress_records_lz77)");
    if (batches.empty()) {                               // empty input: the header alone (block_num = 0)
        DevSlot &s = g_slots[devs[0]];
        std::lock_guard<std::mutex> lk(s.mu);
        int rc = slot_prepare_compress(s, devs[0], 16, 64, 1);
        if (rc) return rc;
        rc = fcx_compress_lz77_device(s.ctx, s.dIn, 0, s.dOut, 64, (withHeader ? FCX_WITH_HEADER : 0) | fmt, hdrBytes, (uint32_t)hdrBlocks);
        uint64_t got = 0;
        if (rc == FCX_OK) rc = fcx_ctx_finish(s.ctx, &got);
        if (rc) return rc;
        if (got > outCap) return fail(FCX_ERR_CAPACITY, "output buffer too small");
        if (got) CK(cudaMemcpy(out, s.dOut, got, cudaMemcpyDeviceToHost));
        *outN = got;
        return FCX_OK;
    }
    Placement place(batches.size(), 0);
    std::vector<int> rcs(G, FCX_OK);
    std::vector<std::string> errs(G);

    auto shard = [&](size_t g) {
        DevSlot &s = g_slots[devs[g]];
        std::lock_guard<std::mutex> lk(s.mu);           // held until this shard's last byte has left the device
        const std::vector<size_t> &my = mine[g];
        if (my.empty()) return;
        int rc = FCX_OK;
        bool secondary = false;                         // gave up only because another shard failed: that one reports
        auto bail = [&](int code) { rcs[g] = code; if (!secondary) errs[g] = last_error_string(); place.fail_all(); };
        uint64_t localBytes = 0;
        uint32_t maxBatch = 1;
        std::vector<uint32_t> sizes;
        for (size_t j : my) {
            const uint64_t b0 = batches[j].firstBlock * (uint64_t)kBlockBytes;
            localBytes += std::min<uint64_t>(n - b0, (uint64_t)batches[j].nBlocks * kBlockBytes);
            maxBatch = std::max(maxBatch, batches[j].nBlocks);
            sizes.push_back(batches[j].nBlocks);
        }
        const uint64_t bound = fcx_compress_bound(localBytes);
        if ((rc = slot_prepare_compress(s, devs[g], localBytes, bound, maxBatch))) return bail(rc);
        // ---- uploads, all queued at once on the copy stream ---------------------------------------------------------
        cudaError_t e = cudaSuccess;
        uint64_t o = 0;
        for (size_t k = 0; k < my.size() && e == cudaSuccess; k++) {
            const GBatch &b = batches[my[k]];
            const uint64_t b0 = b.firstBlock * (uint64_t)kBlockBytes;
            const uint64_t l = std::min<uint64_t>(n - b0, (uint64_t)b.nBlocks * kBlockBytes);
            e = cudaMemcpyAsync(s.dIn + o, in + b0, l, cudaMemcpyHostToDevice, s.h2d);
            if (e == cudaSuccess) e = cudaEventRecord(s.evIn[k], s.h2d);
            o += l;
        }
        if (e != cudaSuccess) rc = fail(FCX_ERR_CUDA, std::string("H2D: ") + cudaGetErrorString(e));
        // ---- kernels of every batch, queued behind their uploads ---------------------------------------------------
        if (rc == FCX_OK) {
            BatchHooks hooks;
            hooks.waitBefore = s.evIn.data(); hooks.recordAfter = s.evOut.data(); hooks.hostCursor = s.hostCursor;
            hooks.sizes = sizes.data(); hooks.nSizes = (uint32_t)sizes.size();
            s.ctx->batchBlocks = 0;
            const bool first = withHeader && my[0] == 0;
            rc = run_batches(s.ctx, s.dIn, localBytes, s.dOut, bound, (first ? FCX_WITH_HEADER : 0) | fmt, hdrBytes, (uint32_t)hdrBlocks, &hooks);
        }
        // ---- downloads: each batch goes to its final place as soon as the batches before it have reported -----------
        uint64_t prev = 0;
        for (size_t k = 0; k < my.size() && rc == FCX_OK; k++) {
            if (cudaEventSynchronize(s.evOut[k]) != cudaSuccess) { rc = fail(FCX_ERR_CUDA, "a device batch failed"); break; }
            const uint64_t cur = s.hostCursor[k];
            if (cur < prev || cur > bound) { rc = fail(FCX_ERR_CAPACITY, "output buffer smaller than the compressed result"); break; }
            uint64_t at = 0;
            if (!place.publish(my[k], cur - prev, at)) { rc = FCX_ERR_CUDA; secondary = true; break; }
            if (at + (cur - prev) > outCap) { rc = fail(FCX_ERR_CAPACITY, "output buffer smaller than the compressed result"); break; }
            if (cur > prev && cudaMemcpyAsync(out + at, s.dOut + prev, cur - prev, cudaMemcpyDeviceToHost, s.d2h) != cudaSuccess)
                rc = fail(FCX_ERR_CUDA, "D2H copy failed");
            prev = cur;
        }
        if (rc != FCX_OK) place.fail_all();
        // nothing of this call may still be in flight when the slot is released, whatever happened above
        uint64_t got = 0;
        const int frc = fcx_ctx_finish(s.ctx, &got);                         // syncs the kernel stream, reports device-side errors
        if (rc == FCX_OK) rc = frc;
        const bool h2dOk = cudaStreamSynchronize(s.h2d) == cudaSuccess;
        const bool d2hOk = cudaStreamSynchronize(s.d2h) == cudaSuccess;
        if (rc == FCX_OK && !(h2dOk && d2hOk)) rc = fail(FCX_ERR_CUDA, "copy stream failed");
        if (rc != FCX_OK) bail(rc);
    };
    if (G == 1) shard(0);
    else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++) th.emplace_back(shard, g);
        for (auto &t : th) t.join();
    }
    // report the first real error (a shard that merely gave up because another failed has no message)
    for (size_t g = 0; g < G; g++) if (rcs[g] && !errs[g].empty()) return fail(rcs[g], errs[g]);
    for (size_t g = 0; g < G; g++) if (rcs[g]) return fail(rcs[g], "a device shard failed");
    *outN = place.total();
    return FCX_OK;
}

} // namespace

extern "C" uint32_t fcx_debug_pipeline_plan(uint64_t blocks, uint32_t *sizes, uint32_t cap)
{
    const std::vector<uint32_t> plan = pipeline_plan(blocks);
    for (size_t i = 0; i < plan.size() && i < cap && sizes; i++) sizes[i] = plan[i];
    return (uint32_t)plan.size();
}

extern "C" int fcx_compress_lz77(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t outCap, uint64_t *outN,
                                 uint32_t deviceMask)
{
    return compress_host(in, n, out, outCap, outN, deviceMask, true, n, (n + kBlockBytes - 1) / kBlockBytes);
}

extern "C" int fcx_compress_records_lz77(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t outCap, uint64_t *outN,
                                         uint32_t deviceMask)
{
    if (n == 0) { if (outN) *outN = 0; return (out && outN) ? FCX_OK : fail(FCX_ERR_ARG, "null pointer"); }
    return compress_host(in, n, out, outCap, outN, deviceMask, false, 0, 0);
}

static int compress_block_host(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t outCap, uint32_t *outN, uint32_t fmt)
{
    if (!in || !out || !outN) return fail(FCX_ERR_ARG, "null pointer");             // mc:2122-2123
    if (n == 0 || n > kBlockBytes) return fail(FCX_ERR_ARG, "a block holds 1..1048576 bytes (mc:4090)");
    if (fcx_device_count() == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    DevSlot &s = g_slots[0];
    std::lock_guard<std::mutex> lk(s.mu);
    const uint64_t bound = fcx_compress_bound(n);
    int rc = slot_prepare_compress(s, 0, n, bound, 1);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s.dIn, in, n, cudaMemcpyHostToDevice, s.stream));
    s.ctx->batchBlocks = 0;
    rc = fcx_compress_lz77_device(s.ctx, s.dIn, n, s.dOut, bound, fmt, 0, 0);
    uint64_t bytes = 0;
    const int frc = fcx_ctx_finish(s.ctx, &bytes);                                  // always: nothing stays in flight
    if (rc == FCX_OK) rc = frc;
    if (rc) return rc;
    if (bytes < 4 || bytes - 4 > outCap) return fail(FCX_ERR_CAPACITY, "output buffer smaller than the payload");
    CK(cudaMemcpyAsync(out, s.dOut + 4, bytes - 4, cudaMemcpyDeviceToHost, s.stream));   // strip the record's u32 size
    CK(cudaStreamSynchronize(s.stream));
    *outN = (uint32_t)(bytes - 4);
    return FCX_OK;
}

extern "C" int fcx_compress_block_lz77(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t outCap, uint32_t *outN)
{
    return compress_block_host(in, n, out, outCap, outN, 0);
}

// ---- FCX9: the opt-in format (64-bit header fields, named single-symbol sections); not readable by the reference ----
extern "C" int fcx_compress_lz77_x9(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t outCap, uint64_t *outN,
                                    uint32_t deviceMask)
{
    const uint64_t nBlocks = (n + kBlockBytes - 1) / kBlockBytes;
    if (nBlocks > 0xFFFFFFFFull) return fail(FCX_ERR_UNSUPPORTED, "more than 2^32 - 1 blocks");
    return compress_host(in, n, out, outCap, outN, deviceMask, true, n, nBlocks, FCX_FORMAT_X9);
}

extern "C" int fcx_compress_records_lz77_x9(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t outCap, uint64_t *outN,
                                            uint32_t deviceMask)
{
    if (n == 0) { if (outN) *outN = 0; return (out && outN) ? FCX_OK : fail(FCX_ERR_ARG, "null pointer"); }
    return compress_host(in, n, out, outCap, outN, deviceMask, false, 0, 0, FCX_FORMAT_X9);
}

extern "C" int fcx_compress_block_lz77_x9(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t outCap, uint32_t *outN)
{
    return compress_block_host(in, n, out, outCap, outN, FCX_FORMAT_X9);
}

// -----------------------------------------------------------------------------------------
// decompress (device kernels in fcx_decode.cu)
// -----------------------------------------------------------------------------------------
namespace {

// records: offsets of the "u32 size | payload" records inside `file`; recEnd = end of the last one
int decompress_host(const uint8_t *file, const std::vector<uint64_t> &rec, uint64_t recEnd, uint8_t *out, uint64_t outCap,
                    uint64_t *outN, uint32_t mask, bool x9 = false)
{
    const int nd = fcx_device_count();
    if (nd == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    const std::vector<int> devs = devices_of(mask, nd);
    if (devs.empty()) return fail(FCX_ERR_ARG, "device_mask selects no visible device");
    const uint64_t nBlocks = rec.size();
    if (nBlocks == 0) { *outN = 0; return FCX_OK; }
    const size_t G = (size_t)std::min<uint64_t>(devs.size(), nBlocks);
    // batches of <= kDecBatchBlocks consecutive blocks, dealt round-robin; small first round so that the first
    // upload is short
    std::vector<GBatch> batches;
    std::vector<std::vector<size_t>> mine(G);
    {
        uint64_t at = 0;
        uint32_t sz = (uint32_t)std::min<uint64_t>(kDecBatchBlocks, std::max<uint64_t>(8, nBlocks / (16 * G)));
        while (at < nBlocks) {
            for (size_t g = 0; g < G && at < nBlocks; g++) {
                const uint32_t nb = (uint32_t)std::min<uint64_t>(sz, nBlocks - at);
                mine[g].push_back(batches.size());
                batches.push_back({at, nb, (uint32_t)g});
                at += nb;
            }
            sz = (uint32_t)std::min<uint64_t>(kDecBatchBlocks, (uint64_t)sz * 4);
        }
    }
    for (size_t g = 0; g < G; g++)
        if (mine[g].size() > kMaxPipeBatches) return fail(FCX_ERR_UNSUPPORTED, "container has too many blocks for one call");
    Placement place(batches.size(), 0);
    std::vector<int> rcs(G, FCX_OK);
    std::vector<std::string> errs(G);
    const size_t per = decode_scratch_per_block();

    auto shard = [&](size_t g) {
        DevSlot &s = g_slots[devs[g]];
        std::lock_guard<std::mutex> lk(s.mu);
        const std::vector<size_t> &my = mine[g];
        if (my.empty()) return;
        int rc = FCX_OK;
        bool secondary = false;
        auto bail = [&](int code) { rcs[g] = code; if (!secondary) errs[g] = last_error_string(); place.fail_all(); };
        auto cuda_ok = [&](cudaError_t e, const char *what) {
            if (e != cudaSuccess && rc == FCX_OK) rc = fail(FCX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
            return e == cudaSuccess;
        };
        // ---- sizes: compressed bytes of my batches (16-byte aligned pieces), blocks, widest batch ----------------
        uint64_t inBytes = 0, myBlocks = 0;
        uint32_t maxBatch = 1;
        for (size_t j : my) {
            const GBatch &b = batches[j];
            const uint64_t f0 = rec[b.firstBlock] & ~15ull;
            const uint64_t f1 = b.firstBlock + b.nBlocks < nBlocks ? rec[b.firstBlock + b.nBlocks] : recEnd;
            inBytes += ((f1 - f0 + 15) & ~15ull) + 16;
            myBlocks += b.nBlocks;
            maxBatch = std::max(maxBatch, b.nBlocks);
        }
        if ((rc = slot_streams(s, devs[g]))) return bail(rc);
        if ((rc = grow(s.dIn, s.inCap, inBytes + 64))) return bail(rc);
        if ((rc = grow(s.dOut, s.outCap, (uint64_t)kDecRing * maxBatch * kBlockBytes + 64))) return bail(rc);
        if ((rc = grow(s.decScratch, s.decScratchCap, (uint64_t)per * maxBatch))) return bail(rc);
        if (s.decBlocks < myBlocks) {
            if (s.decRec) cudaFree(s.decRec);
            if (s.decLen) cudaFree(s.decLen);
            s.decRec = nullptr; s.decLen = nullptr; s.decBlocks = 0;
            if (!cuda_ok(cudaMalloc(&s.decRec, sizeof(uint64_t) * (myBlocks + 1)), "cudaMalloc") ||
                !cuda_ok(cudaMalloc(&s.decLen, sizeof(uint32_t) * (myBlocks + 1)), "cudaMalloc")) return bail(rc);
            s.decBlocks = myBlocks;
        }
        if (s.hostBlocks < myBlocks) {
            if (s.hostRec) cudaFreeHost(s.hostRec);
            if (s.hostLen) cudaFreeHost(s.hostLen);
            s.hostRec = nullptr; s.hostLen = nullptr; s.hostBlocks = 0;
            if (!cuda_ok(cudaHostAlloc(reinterpret_cast<void **>(&s.hostRec), sizeof(uint64_t) * (myBlocks + 1), cudaHostAllocDefault), "cudaHostAlloc") ||
                !cuda_ok(cudaHostAlloc(reinterpret_cast<void **>(&s.hostLen), sizeof(uint32_t) * (myBlocks + 1), cudaHostAllocDefault), "cudaHostAlloc")) return bail(rc);
            s.hostBlocks = myBlocks;
        }
        // ---- uploads of all my batches (copy stream); record offsets relative to the device buffer ---------------
        {
            uint64_t o = 0, blk = 0;
            for (size_t k = 0; k < my.size() && rc == FCX_OK; k++) {
                const GBatch &b = batches[my[k]];
                const uint64_t f0 = rec[b.firstBlock] & ~15ull;
                const uint64_t f1 = b.firstBlock + b.nBlocks < nBlocks ? rec[b.firstBlock + b.nBlocks] : recEnd;
                for (uint32_t i = 0; i < b.nBlocks; i++) s.hostRec[blk + i] = o + (rec[b.firstBlock + i] - f0);
                cuda_ok(cudaMemcpyAsync(s.dIn + o, file + f0, f1 - f0, cudaMemcpyHostToDevice, s.h2d), "H2D");
                // the decoder's bit reader may look a few bytes beyond the last code word (mc:4195 relies on zeroed slack too)
                cuda_ok(cudaMemsetAsync(s.dIn + o + (f1 - f0), 0, (((f1 - f0 + 15) & ~15ull) + 16) - (f1 - f0), s.h2d), "memset");
                cuda_ok(cudaEventRecord(s.evIn[k], s.h2d), "event");
                o += ((f1 - f0 + 15) & ~15ull) + 16;
                blk += b.nBlocks;
            }
            cuda_ok(cudaMemcpyAsync(s.decRec, s.hostRec, sizeof(uint64_t) * myBlocks, cudaMemcpyHostToDevice, s.stream), "H2D");
        }
        // ---- decode batch k into ring slot k % kDecRing; place batch k - kDecRing first, which frees that slot ------
        std::vector<uint64_t> blk0(my.size() + 1, 0);
        for (size_t k = 0; k < my.size(); k++) blk0[k + 1] = blk0[k] + batches[my[k]].nBlocks;
        auto place_batch = [&](size_t k) {
            const GBatch &b = batches[my[k]];
            if (!cuda_ok(cudaEventSynchronize(s.evOut[k]), "decode batch")) return;
            const uint32_t *lens = s.hostLen + blk0[k];
            uint64_t total = 0;
            bool packed = true;                                    // every block but the last is full: one copy
            for (uint32_t i = 0; i < b.nBlocks; i++) {
                if (lens[i] == 0xFFFFFFFFu) { rc = fail(FCX_ERR_FORMAT, "corrupt block payload"); return; }
                if (i + 1 < b.nBlocks && lens[i] != kBlockBytes) packed = false;
                total += lens[i];
            }
            uint64_t at = 0;
            if (!place.publish(my[k], total, at)) { rc = FCX_ERR_CUDA; secondary = true; return; }
            if (at + total > outCap) { rc = fail(FCX_ERR_CAPACITY, "output buffer too small"); return; }
            const uint8_t *src = s.dOut + (uint64_t)(k % kDecRing) * maxBatch * kBlockBytes;
            if (packed) {
                if (total) cuda_ok(cudaMemcpyAsync(out + at, src, total, cudaMemcpyDeviceToHost, s.d2h), "D2H");
            } else {
                for (uint32_t i = 0; i < b.nBlocks; i++) {
                    if (lens[i]) cuda_ok(cudaMemcpyAsync(out + at, src + (uint64_t)i * kBlockBytes, lens[i], cudaMemcpyDeviceToHost, s.d2h), "D2H");
                    at += lens[i];
                }
            }
            cuda_ok(cudaEventRecord(s.evFree[k], s.d2h), "event");
        };
        for (size_t k = 0; k < my.size() && rc == FCX_OK; k++) {
            if (k >= kDecRing) {
                place_batch(k - kDecRing);
                if (rc != FCX_OK) break;
                cuda_ok(cudaStreamWaitEvent(s.stream, s.evFree[k - kDecRing], 0), "wait");
            }
            const GBatch &b = batches[my[k]];
            cuda_ok(cudaStreamWaitEvent(s.stream, s.evIn[k], 0), "wait");
            if (launch_decode(s.dIn, s.decRec + blk0[k], b.nBlocks, s.dOut + (uint64_t)(k % kDecRing) * maxBatch * kBlockBytes,
                              s.decLen + blk0[k], s.decScratch, per, s.stream, x9) < 0) {
                rc = fail(FCX_ERR_CUDA, "decode kernels failed to launch");
                break;
            }
            cuda_ok(cudaMemcpyAsync(s.hostLen + blk0[k], s.decLen + blk0[k], sizeof(uint32_t) * b.nBlocks, cudaMemcpyDeviceToHost, s.stream), "D2H");
            cuda_ok(cudaEventRecord(s.evOut[k], s.stream), "event");
        }
        for (size_t k = my.size() > kDecRing ? my.size() - kDecRing : 0; k < my.size() && rc == FCX_OK; k++) place_batch(k);
        if (rc != FCX_OK) place.fail_all();
        const bool a = cudaStreamSynchronize(s.stream) == cudaSuccess, b2 = cudaStreamSynchronize(s.h2d) == cudaSuccess,
                   c2 = cudaStreamSynchronize(s.d2h) == cudaSuccess;
        if (rc == FCX_OK && !(a && b2 && c2)) rc = fail(FCX_ERR_CUDA, std::string("decode pipeline: ") + cudaGetErrorString(cudaGetLastError()));
        if (rc != FCX_OK) bail(rc);
    };
    if (G == 1) shard(0);
    else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++) th.emplace_back(shard, g);
        for (auto &t : th) t.join();
    }
    for (size_t g = 0; g < G; g++) if (rcs[g] && !errs[g].empty()) return fail(rcs[g], errs[g]);
    for (size_t g = 0; g < G; g++) if (rcs[g]) return fail(rcs[g], "a device shard failed");
    *outN = place.total();
    return FCX_OK;
}

// walks `count` records starting at file[pos] (mc:4162-4174)
int scan_records(const uint8_t *file, uint64_t nbytes, uint64_t pos, uint64_t count, std::vector<uint64_t> &rec, uint64_t &end)
{
    rec.reserve(count);
    for (uint64_t b = 0; b < count; b++) {
        if (pos + 4 > nbytes) return fail(FCX_ERR_FORMAT, "truncated record header");
        uint32_t sz; memcpy(&sz, file + pos, 4);
        if (pos + 4 + (uint64_t)sz > nbytes) return fail(FCX_ERR_FORMAT, "truncated record");
        rec.push_back(pos);
        pos += 4 + (uint64_t)sz;
    }
    end = pos;
    return FCX_OK;
}

} // namespace

namespace {
int decompress_lz78_host(const uint8_t *file, const std::vector<uint64_t> &rec, uint64_t recEnd, uint8_t *out, uint64_t outCap,
                         uint64_t *outN, uint32_t mask);
}

extern "C" int fcx_decompress_lz77(const uint8_t *file, uint64_t nbytes, uint8_t *out, uint64_t outCap, uint64_t *outN,
                                   uint32_t deviceMask)
{
    if (!file || !out || !outN) return fail(FCX_ERR_ARG, "null pointer");
    if (nbytes < 10) return fail(FCX_ERR_FORMAT, "shorter than the 10-byte header");               // mc:4139-4144
    if (memcmp(file, "FCX", 3) != 0) return fail(FCX_ERR_FORMAT, "not an FCX container");          // mc:4146-4152
    if (file[3] != '7' && file[3] != '8' && file[3] != '9') return fail(FCX_ERR_FORMAT, "not an FCX7 / FCX8 / FCX9 container");   // mc:4153-4164
    if (fcx_device_count() == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    std::vector<uint64_t> rec;
    if (file[3] == '9') {                                  // opt-in format: u64 byte count, u32 block count
        if (nbytes < FCX_X9_HEADER_BYTES) return fail(FCX_ERR_FORMAT, "shorter than the 16-byte FCX9 header");
        uint64_t total; uint32_t nb32;
        memcpy(&total, file + 4, 8); memcpy(&nb32, file + 12, 4);
        if ((total + kBlockBytes - 1) / kBlockBytes != nb32) return fail(FCX_ERR_FORMAT, "FCX9 header: block count does not match the byte count");
        if ((uint64_t)nb32 * 4 > nbytes) return fail(FCX_ERR_FORMAT, "truncated FCX9 container");
        uint64_t end9 = FCX_X9_HEADER_BYTES;
        int rc9 = scan_records(file, nbytes, FCX_X9_HEADER_BYTES, nb32, rec, end9);
        if (rc9) return rc9;
        rc9 = decompress_host(file, rec, end9, out, outCap, outN, deviceMask, true);
        if (rc9 == FCX_OK && *outN != total) return fail(FCX_ERR_FORMAT, "FCX9: decoded byte count differs from the header");
        return rc9;
    }
    uint16_t nb16; memcpy(&nb16, file + 8, 2);
    uint64_t end = 10;
    int rc = scan_records(file, nbytes, 10, nb16, rec, end);
    if (rc) return rc;
    if (file[3] == '8') return decompress_lz78_host(file, rec, end, out, outCap, outN, deviceMask);   // mc:4189
    return decompress_host(file, rec, end, out, outCap, outN, deviceMask);
}

extern "C" int fcx_decompress_records_lz77(const uint8_t *records, uint64_t nbytes, uint32_t nBlocks, uint8_t *out,
                                           uint64_t outCap, uint64_t *outN, uint32_t deviceMask)
{
    if (!records || !out || !outN) return fail(FCX_ERR_ARG, "null pointer");
    if (fcx_device_count() == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    std::vector<uint64_t> rec;
    uint64_t end = 0;
    int rc = scan_records(records, nbytes, 0, nBlocks, rec, end);
    if (rc) return rc;
    return decompress_host(records, rec, end, out, outCap, outN, deviceMask);
}

static int decompress_block_host(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t outCap, uint32_t *outN, bool x9);

extern "C" int fcx_decompress_block_lz77(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t outCap,
                                         uint32_t *outN)
{
    return decompress_block_host(payload, nbytes, out, outCap, outN, false);
}

extern "C" int fcx_decompress_block_lz77_x9(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t outCap,
                                            uint32_t *outN)
{
    return decompress_block_host(payload, nbytes, out, outCap, outN, true);
}

extern "C" int fcx_decompress_records_lz77_x9(const uint8_t *records, uint64_t nbytes, uint32_t nBlocks, uint8_t *out,
                                              uint64_t outCap, uint64_t *outN, uint32_t deviceMask)
{
    if (!records || !out || !outN) return fail(FCX_ERR_ARG, "null pointer");
    if (fcx_device_count() == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    std::vector<uint64_t> rec;
    uint64_t end = 0;
    int rc = scan_records(records, nbytes, 0, nBlocks, rec, end);
    if (rc) return rc;
    return decompress_host(records, rec, end, out, outCap, outN, deviceMask, true);
}

static int decompress_block_host(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t outCap, uint32_t *outN, bool x9)
{
    if (!payload || !out || !outN) return fail(FCX_ERR_ARG, "null pointer");                      // mc:2257-2258
    if (fcx_device_count() == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    std::vector<uint8_t> rec(4 + (size_t)nbytes);
    memcpy(rec.data(), &nbytes, 4);
    memcpy(rec.data() + 4, payload, nbytes);
    std::vector<uint64_t> off(1, 0);
    uint64_t got = 0;
    int rc = decompress_host(rec.data(), off, rec.size(), out, outCap, &got, 1, x9);
    if (rc) return rc;
    *outN = (uint32_t)got;
    return FCX_OK;
}

// -----------------------------------------------------------------------------------------
// FCX8 (lz78): "-c <anything but lz77>" of the reference CLI (mc:4040-4043, mc:4106)
// -----------------------------------------------------------------------------------------
// The tokenizer walks every block sequentially (one thread per block, see fcx_lz78.cu), so the device time
// dwarfs the copies: one upload, the kernels, one download, on the first device of the mask.
namespace {
int compress_lz78_host(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t outCap, uint64_t *outN, uint32_t mask, bool withHeader)
{
    if (!out || !outN || (!in && n)) return fail(FCX_ERR_ARG, "null pointer");
    const int nd = fcx_device_count();
    if (nd == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    const std::vector<int> devs = devices_of(mask, nd);
    if (devs.empty()) return fail(FCX_ERR_ARG, "device_mask selects no visible device");
    DevSlot &s = g_slots[devs[0]];
    std::lock_guard<std::mutex> lk(s.mu);
    const uint64_t nBlocks = (n + kBlockBytes - 1) / kBlockBytes;
    // lz78 payloads can exceed the lz77 bound (32-bit tree nodes, group codes): the reference's own buffer is 2 MiB per block
    const uint64_t bound = 10 + nBlocks * (4 + (5ull << 19) + (1ull << 20) + 4096) + 64;
    int rc = slot_prepare_compress(s, devs[0], std::max<uint64_t>(n, 16), bound, (uint32_t)std::min<uint64_t>(std::max<uint64_t>(nBlocks, 1), kMaxBatchBlocks));
    if (rc) return rc;
    if (n) CK(cudaMemcpyAsync(s.dIn, in, n, cudaMemcpyHostToDevice, s.stream));
    uint64_t got = 0;
    rc = lz78_compress_device(s.ctx, s.dIn, n, s.dOut, bound, withHeader, n, (uint32_t)nBlocks, &got);
    if (rc) { cudaStreamSynchronize(s.stream); return rc; }
    if (got > outCap) return fail(FCX_ERR_CAPACITY, "output buffer smaller than the compressed result");
    if (got) CK(cudaMemcpyAsync(out, s.dOut, got, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    *outN = got;
    return FCX_OK;
}
} // namespace

namespace {
// FCX8 records -> bytes: upload, decode in batches of <= 64 blocks (15 MB of scratch each), one download per block
int decompress_lz78_host(const uint8_t *file, const std::vector<uint64_t> &rec, uint64_t recEnd, uint8_t *out, uint64_t outCap,
                         uint64_t *outN, uint32_t mask)
{
    const int nd = fcx_device_count();
    if (nd == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    const std::vector<int> devs = devices_of(mask, nd);
    if (devs.empty()) return fail(FCX_ERR_ARG, "device_mask selects no visible device");
    if (rec.empty()) { *outN = 0; return FCX_OK; }
    DevSlot &s = g_slots[devs[0]];
    std::lock_guard<std::mutex> lk(s.mu);
    int rc = slot_streams(s, devs[0]);
    if (rc) return rc;
    lz78_decode_init();
    constexpr uint32_t kBatch = 64;
    constexpr size_t kStride = (size_t)kBlockBytes + 256;
    const size_t per = lz78_decode_scratch_per_block();
    const uint64_t f0 = rec[0] & ~15ull, bytes = recEnd - f0;
    if ((rc = grow(s.dIn, s.inCap, bytes + 64))) return rc;
    if ((rc = grow(s.dOut, s.outCap, kBatch * kStride))) return rc;
    if ((rc = grow(s.decScratch, s.decScratchCap, per * kBatch))) return rc;
    uint64_t *dRec = nullptr; uint32_t *dLen = nullptr;
    CK(cudaMalloc(&dRec, sizeof(uint64_t) * kBatch));
    CK(cudaMalloc(&dLen, sizeof(uint32_t) * kBatch));
    auto done = [&](int code) { cudaStreamSynchronize(s.stream); cudaFree(dRec); cudaFree(dLen); return code; };
    if (cudaMemcpyAsync(s.dIn, file + f0, bytes, cudaMemcpyHostToDevice, s.stream) != cudaSuccess ||
        cudaMemsetAsync(s.dIn + bytes, 0, 32, s.stream) != cudaSuccess) return done(fail(FCX_ERR_CUDA, "H2D copy failed"));
    uint64_t at = 0;
    for (size_t b0 = 0; b0 < rec.size(); b0 += kBatch) {
        const uint32_t nb = (uint32_t)std::min<size_t>(kBatch, rec.size() - b0);
        uint64_t offs[kBatch];
        uint32_t lens[kBatch];
        for (uint32_t i = 0; i < nb; i++) offs[i] = rec[b0 + i] - f0;
        if (cudaMemcpyAsync(dRec, offs, sizeof(uint64_t) * nb, cudaMemcpyHostToDevice, s.stream) != cudaSuccess)
            return done(fail(FCX_ERR_CUDA, "H2D copy failed"));
        if (launch_decode_lz78(s.dIn, dRec, nb, s.dOut, kStride, dLen, s.decScratch, per, s.stream) < 0)
            return done(fail(FCX_ERR_CUDA, "lz78 decode kernel failed to launch"));
        if (cudaMemcpyAsync(lens, dLen, sizeof(uint32_t) * nb, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess ||
            cudaStreamSynchronize(s.stream) != cudaSuccess)
            return done(fail(FCX_ERR_CUDA, std::string("lz78 decode: ") + cudaGetErrorString(cudaGetLastError())));
        for (uint32_t i = 0; i < nb; i++) {
            if (lens[i] == 0xFFFFFFFFu) return done(fail(FCX_ERR_FORMAT, "corrupt block payload"));
            if (at + lens[i] > outCap) return done(fail(FCX_ERR_CAPACITY, "output buffer too small"));
            if (lens[i] && cudaMemcpyAsync(out + at, s.dOut + (size_t)i * kStride, lens[i], cudaMemcpyDeviceToHost, s.stream) != cudaSuccess)
                return done(fail(FCX_ERR_CUDA, "D2H copy failed"));
            at += lens[i];
        }
        if (cudaStreamSynchronize(s.stream) != cudaSuccess) return done(fail(FCX_ERR_CUDA, "D2H copy failed"));
    }
    *outN = at;
    return done(FCX_OK);
}
} // namespace

extern "C" int fcx_decompress_block_lz78(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t outCap, uint32_t *outN)
{
    if (!payload || !out || !outN) return fail(FCX_ERR_ARG, "null pointer");
    if (fcx_device_count() == 0) return fail(FCX_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    std::vector<uint8_t> rec(4 + (size_t)nbytes);
    memcpy(rec.data(), &nbytes, 4);
    memcpy(rec.data() + 4, payload, nbytes);
    std::vector<uint64_t> off(1, 0);
    uint64_t got = 0;
    const int rc = decompress_lz78_host(rec.data(), off, rec.size(), out, outCap, &got, 1);
    if (rc) return rc;
    *outN = (uint32_t)got;
    return FCX_OK;
}

extern "C" uint64_t fcx_compress_bound_lz78(uint64_t n)
{
    const uint64_t nBlocks = (n + kBlockBytes - 1) / kBlockBytes;
    return 10 + nBlocks * (4 + (5ull << 19) + (1ull << 20) + 4096) + 64;
}

extern "C" int fcx_compress_lz78(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t outCap, uint64_t *outN, uint32_t deviceMask)
{
    return compress_lz78_host(in, n, out, outCap, outN, deviceMask, true);
}

extern "C" int fcx_compress_block_lz78(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t outCap, uint32_t *outN)
{
    if (!in || !out || !outN) return fail(FCX_ERR_ARG, "null pointer");
    if (n == 0 || n > kBlockBytes) return fail(FCX_ERR_ARG, "a block holds 1..1048576 bytes (mc:4090)");
    std::vector<uint8_t> rec((size_t)fcx_compress_bound_lz78(n));
    uint64_t got = 0;
    const int rc = compress_lz78_host(in, n, rec.data(), rec.size(), &got, 1, false);
    if (rc) return rc;
    if (got < 4 || got - 4 > outCap) return fail(FCX_ERR_CAPACITY, "output buffer smaller than the payload");
    memcpy(out, rec.data() + 4, got - 4);                  // strip the record's u32 size
    *outN = (uint32_t)(got - 4);
    return FCX_OK;
}

// -----------------------------------------------------------------------------------------
// pinned host memory for callers (file readers that want the copies to run at PCIe speed)
// -----------------------------------------------------------------------------------------
extern "C" void *fcx_host_alloc(uint64_t bytes)
{
    void *p = nullptr;
    if (fcx_device_count() == 0 || cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        fail(FCX_ERR_CUDA, "pinned allocation failed");
        return nullptr;
    }
    return p;
}

extern "C" void fcx_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}
