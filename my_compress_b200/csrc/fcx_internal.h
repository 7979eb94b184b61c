// fcx_internal.h -- host-side declarations shared by fcx_api.cu (context, device-resident path, probes)
// and fcx_host.cu (host-buffer pipelines, multi-GPU sharding).  Not part of the public ABI.
#pragma once
#include "../../include/fcx.h"
#include "fcx_dev.cuh"

#include <string>
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
int launch_huff_stage(const Ws &ws, uint32_t nBlocks, uint8_t *out, uint64_t outCap, uint32_t flags,   // FCX_WITH_HEADER | FCX_FORMAT_X9
                      uint64_t hdrTotalBytes, uint32_t hdrBlocks, uint64_t batchStartCursor, int numSM,
                      cudaStream_t st, void (*mark)(void *, int), void *markArg);
// fcx_decode.cu
int launch_decode(const uint8_t *d_file, const uint64_t *d_recOff, uint32_t nBlocks, uint8_t *d_out,
                  uint32_t *d_outLen, uint8_t *d_scratch, size_t scratchPerBlock, cudaStream_t st, bool x9 = false);
size_t decode_scratch_per_block();
void decode_init();

// fcx_lz78.cu
} // namespace fcx
struct fcx_ctx;
namespace fcx {
int lz78_compress_device(fcx_ctx *c, const uint8_t *d_in, uint64_t n, uint8_t *d_out, uint64_t outCap, bool withHeader,
                         uint64_t hdrBytes, uint32_t hdrBlocks, uint64_t *outBytes);
void lz78_release(fcx_ctx *c);
size_t lz78_decode_scratch_per_block();
void lz78_decode_init();
int launch_decode_lz78(const uint8_t *d_file, const uint64_t *d_recOff, uint32_t nBlocks, uint8_t *d_out, size_t outStride,
                       uint32_t *d_outLen, uint8_t *d_scratch, size_t scratchPerBlock, cudaStream_t st);

// ---- errors (thread-local message behind fcx_last_error) --------------------------------
int fail(int code, const std::string &msg);
const std::string &last_error_string();
} // namespace fcx

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fcx::fail(FCX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

enum { ST_MATCH = 0, ST_PARSE, ST_PLAN, ST_HIST, ST_TREE, ST_LAYOUT, ST_ENCODE, ST_ENCODE_SLOW, ST_HEADERS, ST_COUNT };

// The context of the device-resident path (include/fcx.h: fcx_ctx_*).
struct fcx_ctx {
    int device = 0;
    int numSM = 148;
    cudaStream_t stream = 0;
    uint32_t maxBlocks = 0;
    uint32_t batchBlocks = 0;          // blocks per device batch (<= maxBlocks); 0 = maxBlocks
    fcx::Ws ws{};
    std::vector<void *> allocs;
    std::vector<uint8_t *> guardAt;    // 256-byte guard bands before and after every workspace array
    uint64_t launches = 0;
    // pinned mirror of the device counters + totals (one async read per finish)
    uint64_t *hostStatus = nullptr;
    // timing
    bool timing = false;
    std::vector<cudaEvent_t> ev;       // per call: ST_COUNT+1 events per batch
    std::vector<int> evStage;
    float stageMs[ST_COUNT] = {0};
    // device-resident decode scratch (allocated on first use)
    uint8_t *decScratch = nullptr;
    uint64_t *decRec = nullptr;
    uint32_t *decLen = nullptr;
    void *lz78 = nullptr;              // FCX8 workspace (fcx_lz78.cu), allocated on first use
    // last call
    bool pending = false;
    uint64_t lastBytes = 0;
};

namespace fcx {
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
int run_batches(fcx_ctx *c, const uint8_t *d_in, uint64_t n, uint8_t *d_out, uint64_t outCap, uint32_t flags,
                uint64_t hdrBytes, uint32_t hdrBlocks, const BatchHooks *hooks = nullptr);
} // namespace fcx
