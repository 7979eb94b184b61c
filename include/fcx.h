/*
 * fcx.h -- C ABI of the B200-native drop-in for the "-c lz77" path of my_compress.cpp.
 *
 * Every entry point below replaces one reference interface ("mc:N" = line N of
 * /root/reference/my_compress.cpp, \n-based).  Plain pointers and sizes only; errors are
 * negative return codes (the reference exit(-1)s on allocation failure, mc:73-93, and
 * returns 0 on NULL arguments, mc:2122-2123 -- a library must not do either silently).
 * All compute runs in hand-written sm_100a CUDA kernels; there is no CPU fallback: without
 * a CUDA device every compute call returns FCX_ERR_NO_DEVICE.
 *
 * Output is byte-identical to the reference: FCX7 container (mc:101-109, 4073-4136) of
 * independent 1 MiB blocks, each "u32 size | payload" with the payload of mc:2115-2253.
 */
#ifndef FCX_H
#define FCX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCX_BLOCK_BYTES (1u << 20) /* mc:113 BLOCK_BYTES */
#define FCX_HEADER_BYTES 10u       /* mc:101-109 stCmpFileHead, #pragma pack(1) */
#define FCX_X9_HEADER_BYTES 16u    /* the opt-in FCX9 container below */

enum {
    FCX_OK = 0,
    FCX_ERR_ARG = -1,       /* NULL pointer / bad size / misaligned device pointer */
    FCX_ERR_NO_DEVICE = -2, /* no CUDA device, or the kernels could not be loaded */
    FCX_ERR_CUDA = -3,      /* a CUDA runtime call failed; see fcx_last_error() */
    FCX_ERR_CAPACITY = -4,  /* output buffer smaller than the compressed result */
    FCX_ERR_FORMAT = -5,    /* not an FCX7 stream / corrupt payload (mc:4146-4152) */
    FCX_ERR_UNSUPPORTED = -6 /* e.g. FCX8 (lz78) containers: not on this path */
};

/* Thread-local description of the last error on the calling thread ("" if none). */
const char *fcx_last_error(void);

/* Number of CUDA devices visible (0 when none; never an error). */
int fcx_device_count(void);

/* Worst-case container size for n input bytes (the reference's caller uses a fixed 2 MiB
 * per-block buffer, mc:4088; this is the tight version of that convention). */
uint64_t fcx_compress_bound(uint64_t n);

/* ---- host-buffer entry points (H2D/D2H inside) ------------------------------------ */

/* Replaces my_compress_file_lz77(void *pData, uInt32 totalBytes, uInt8 *pOut) (mc:2115,
 * called at mc:4099): one block (n <= 1 MiB) -> payload bytes, written to out, count in
 * *out_n.  Does not print (the reference printf()s per block, mc:2249). */
int fcx_compress_block_lz77(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t out_cap, uint32_t *out_n);

/* Replaces the compress branch of main() (mc:4073-4136): whole buffer -> FCX7 container,
 * header included (u32 byte count and u16 block count wrap exactly like the reference).
 * device_mask: bit g set = use CUDA device g; 0 = device 0.  Blocks are sharded over the
 * devices in contiguous ranges; no data-path collective (blocks share no state). */
int fcx_compress_lz77(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_n,
                      uint32_t device_mask);

/* The body of the block loop alone (mc:4090-4122) for a caller that streams a file through in pieces and
 * writes the header itself, as the reference does (placeholder at mc:4078-4086, rewritten at mc:4128-4129):
 * n bytes (every piece but the last a multiple of 1 MiB) -> their "u32 size | payload" records, no header.
 * Concatenating the records of consecutive pieces behind a header gives the same file as one call. */
int fcx_compress_records_lz77(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_n,
                              uint32_t device_mask);

/* Replaces my_decompress_file_lz77 (mc:2255, called at mc:4182) for one payload, and the
 * decompress branch of main() (mc:4138-4202) for a container.  Output is what the
 * reference's decoder produces, including its single-symbol-stream behaviour. */
int fcx_decompress_block_lz77(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t out_cap,
                              uint32_t *out_n);
int fcx_decompress_lz77(const uint8_t *file, uint64_t nbytes, uint8_t *out, uint64_t out_cap, uint64_t *out_n,
                        uint32_t device_mask);
/* The record loop alone (mc:4162-4197): n_blocks consecutive "u32 size | payload" records, no header. */
int fcx_decompress_records_lz77(const uint8_t *records, uint64_t nbytes, uint32_t n_blocks, uint8_t *out,
                                uint64_t out_cap, uint64_t *out_n, uint32_t device_mask);

/* The lz78 twins: "-c <anything but lz77>" of the reference CLI writes an FCX8 container (mc:4040-4043, 4080-4081)
 * with my_compress_file_lz78 (mc:3127, called at mc:4106) per block.  Same bytes as the reference; payloads can be
 * larger than the input (fcx_compress_bound_lz78).  fcx_decompress_lz77 / fcx_decompress_block_lz78 read them back
 * (my_decompress_file_lz78, mc:3478, called at mc:4189), including the unconditional drop of a trailing 0x00
 * (mc:3701-3703). */
uint64_t fcx_compress_bound_lz78(uint64_t n);
int fcx_compress_block_lz78(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t out_cap, uint32_t *out_n);
int fcx_compress_lz78(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_n, uint32_t device_mask);
int fcx_decompress_block_lz78(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t out_cap, uint32_t *out_n);

/* ---- FCX9: opt-in format fixes -- NOT drop-in, the reference cannot read these files --------------
 * The reference's container has two defects a caller may want to be rid of: the header holds a u32 byte count and a
 * u16 block count (mc:101-109: an 8 GiB input wraps to 0, more than 65 535 blocks are unrepresentable), and a Huffman
 * section whose stream holds one distinct symbol drops that symbol (mc:1013-1019), so such blocks decode to zeros
 * (5000 x 'A', a 1-byte file, 64 distinct bytes).  FCX9 changes exactly those two things:
 *   header   'F','C','X','9' | u64 total_bytes | u32 n_blocks                  (16 bytes, little-endian)
 *   records  u32 size | payload, the payload of mc:2115-2253 except that a section with no tree nodes is
 *            00 | symbol | u32 W = 0  (6 bytes instead of 5) and decodes to `symbol` repeated
 * Every other byte is what the FCX7 path writes, so a block without a single-symbol stream has the same payload in
 * both formats; decode(encode(x)) == x for every x.  fcx_compress_bound() covers FCX9 as well.
 * fcx_decompress_lz77 recognises the magic; the block / record forms need the _x9 twins (a payload has no magic). */
int fcx_compress_lz77_x9(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_n, uint32_t device_mask);
int fcx_compress_records_lz77_x9(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_n,
                                 uint32_t device_mask);
int fcx_compress_block_lz77_x9(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t out_cap, uint32_t *out_n);
int fcx_decompress_block_lz77_x9(const uint8_t *payload, uint32_t nbytes, uint8_t *out, uint32_t out_cap, uint32_t *out_n);
int fcx_decompress_records_lz77_x9(const uint8_t *records, uint64_t nbytes, uint32_t n_blocks, uint8_t *out,
                                   uint64_t out_cap, uint64_t *out_n, uint32_t device_mask);

/* Host buffers: any memory works, page-locked memory lets the copies run at full PCIe speed and overlap the
 * kernels (the reference's fread/fwrite buffers, mc:4073 and mc:4088, are what these replace).  NULL on
 * failure.  Pageable buffers are copied through the driver's staging area: correct, but slower. */
void *fcx_host_alloc(uint64_t bytes);
void fcx_host_free(void *p);

/* ---- device-resident entry points (what the throughput metric times) --------------- */

typedef struct fcx_ctx fcx_ctx;

/* Workspace for batches of up to max_batch_blocks (1..1024) 1 MiB blocks on CUDA device `device`
 * (about 4.7 MiB of HBM per block); longer inputs run as several batches on the same workspace.
 * stream: a cudaStream_t (0 = legacy default). */
int fcx_ctx_create(fcx_ctx **ctx, int device, uint32_t max_batch_blocks);
void fcx_ctx_destroy(fcx_ctx *ctx);
int fcx_ctx_set_stream(fcx_ctx *ctx, void *cuda_stream);

/* Compress n bytes resident at d_in (16-byte aligned device pointer) into d_out
 * (16-byte aligned, out_cap bytes, >= fcx_compress_bound(n) is always enough).
 * flags & FCX_WITH_HEADER: emit the 10-byte file header first, describing
 * (total_bytes_for_header, total_blocks_for_header) -- a rank compressing one shard of a
 * larger file passes the whole file's figures on shard 0 and no header elsewhere.
 * Asynchronous on the context's stream; fcx_ctx_finish() synchronises and reports. */
#define FCX_WITH_HEADER 1u
#define FCX_FORMAT_X9 2u /* write FCX9 records (and, with FCX_WITH_HEADER, its 16-byte header) instead of FCX7 */
int fcx_compress_lz77_device(fcx_ctx *ctx, const uint8_t *d_in, uint64_t n, uint8_t *d_out, uint64_t out_cap,
                             uint32_t flags, uint64_t total_bytes_for_header, uint32_t total_blocks_for_header);
/* Waits for the stream; *out_n = bytes written by the last fcx_compress_lz77_device call. */
int fcx_ctx_finish(fcx_ctx *ctx, uint64_t *out_n);

/* Device-resident twin of fcx_decompress_lz77's record loop (mc:4162-4197): the container (or any run of
 * records) lies at d_file, rec_off[b] (host array) is the byte offset of block b's "u32 size | payload" record
 * inside it, block b decodes to d_out + b MiB and its decoded length is returned in lens[b] (host array).
 * d_file must be readable for 16 bytes beyond the last record.  Synchronous.  The _x9 form reads FCX9 records. */
int fcx_decompress_lz77_device(fcx_ctx *ctx, const uint8_t *d_file, const uint64_t *rec_off, uint32_t n_blocks,
                               uint8_t *d_out, uint32_t *lens);
int fcx_decompress_lz77_device_x9(fcx_ctx *ctx, const uint8_t *d_file, const uint64_t *rec_off, uint32_t n_blocks,
                                  uint8_t *d_out, uint32_t *lens);

/* Kernel launches issued on this context since creation (bench.py's gpu_launches). */
uint64_t fcx_ctx_launch_count(const fcx_ctx *ctx);

/* Per-stage device figures of the last compress call, for the roofline report:
 * s_in  = bytes fed to the Huffman encode kernel (all four streams of all blocks),
 * s_out = bytes of code words it wrote, tokens / matches = LZ77 totals. */
typedef struct {
    uint64_t s_in, s_out, tokens, matches, container_bytes, blocks;
} fcx_stats;
int fcx_ctx_stats(fcx_ctx *ctx, fcx_stats *st);

/* Per-kernel timing hooks: when enabled, each stage of the next compress calls is
 * bracketed by CUDA events on the context's stream.  ms[i] accumulates per stage,
 * names[i] are static strings; returns the number of stages. */
int fcx_ctx_enable_timing(fcx_ctx *ctx, int on);
int fcx_ctx_stage_times(fcx_ctx *ctx, const char **names, float *ms, int cap);

/* ---- stage-level probes (device pointers; used by tests/ to localise parity) -------- */

/* per-position longest match of one block (mc:1446-1514): len[i] in {0,3..257}, dist[i].
 * fcx_probe_match evaluates EVERY position; fcx_probe_match_auto runs the kernel as the product
 * path does, where a tile of long matches only evaluates the positions a parse can reach
 * (every position on the parse chain is exact, unreachable ones read 0). */
int fcx_probe_match(fcx_ctx *ctx, const uint8_t *d_in, uint32_t n, uint16_t *d_len, uint16_t *d_dist);
int fcx_probe_match_auto(fcx_ctx *ctx, const uint8_t *d_in, uint32_t n, uint16_t *d_len, uint16_t *d_dist);
/* greedy parse + stream split of one block (mc:1675-1714, 2073-2113, 2150-2161, 1292-1313,
 * 258-304): host outputs.  streams: bitmap bytes, c bytes, packed-p bytes, Golomb bytes. */
int fcx_probe_streams(fcx_ctx *ctx, const uint8_t *d_in, uint32_t n, uint32_t *T, uint32_t *Pc, uint32_t *G,
                      uint8_t *bitmap, uint8_t *c, uint8_t *ppack, uint8_t *golomb);
/* one Huffman section (mc:987-1104) of S device-resident bytes -> host bytes. */
int fcx_probe_huffman_section(fcx_ctx *ctx, const uint8_t *d_src, uint32_t S, uint8_t *out, uint32_t out_cap,
                              uint32_t *out_n);

/* Memory-safety probe for tests (no sanitizer on the GPU pool): every workspace array of a context lies between
 * two 256-byte guard bands.  check = 0 paints them; check = 1 returns how many bands were written to since. */
int fcx_ctx_debug_guards(fcx_ctx *ctx, int check);

/* The host path's pipeline plan for one device (tests): how `blocks` 1 MiB blocks are cut into device batches
 * (a short first and last batch so that only a small upload / download is exposed, large ones between).
 * Writes up to cap batch sizes to sizes[] and returns the number of batches. */
uint32_t fcx_debug_pipeline_plan(uint64_t blocks, uint32_t *sizes, uint32_t cap);

/* Corpus generator following RandM() of the reference's generator program
 * (随机数的生成.cpp:26-42) with time() pinned to `seed`: byte i = perm[i] & 0xFF. */
void fcx_datagen_randm(uint8_t *out, uint32_t m, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif
