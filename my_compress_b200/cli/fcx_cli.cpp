// fcx_cli -- flag-compatible stand-in for the reference CLI (SURVEY 8f "next" row 2):
//     ./fcx_cli -i <in> [-o <out>] [-c lz77|fcx9] [-g <gpus>]     (mc:3992-4053: no -c => decompress, default out "./out")
//
// Same file loop as the reference's main(), with the three steps of every iteration running side by side:
//   compress   (mc:4073-4136)  placeholder header, then per piece  fread -> fcx_compress_records_lz77 -> fwrite,
//                              header rewritten at the end (mc:4128-4129)
//   decompress (mc:4138-4202)  header, then per piece  fread of whole records -> fcx_decompress_records_lz77 -> fwrite
// A piece is 256 MiB of input (256 blocks; the reference's piece is one block).  A reader thread fills one
// page-locked buffer while the GPU works on the other and a writer thread drains the previous result, so disk,
// PCIe and kernels overlap; the bytes written are identical to the reference's.  Unlike the reference it prints
// nothing but errors.  "-c <anything else>" means lz78 as in the reference (mc:4040-4043): FCX8 files go through
// fcx_compress_lz78 / fcx_decompress_lz77 in one piece (the lz78 tokenizer is a sequential walk per block; there
// is nothing for a file pipeline to overlap with).
#include "../../include/fcx.h"

#include <getopt.h>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

namespace {

#ifndef FCX_CLI_PIECE_BLOCKS
#define FCX_CLI_PIECE_BLOCKS 256         // tests build a variant with tiny pieces to walk the multi-piece path on small files
#endif
constexpr uint64_t kPieceBlocks = FCX_CLI_PIECE_BLOCKS;
constexpr uint64_t kPieceBytes = kPieceBlocks * FCX_BLOCK_BYTES;

// A slot travels reader -> worker -> writer -> reader.
struct Slot {
    uint8_t *in = nullptr, *out = nullptr;
    uint64_t inBytes = 0, outBytes = 0;
    uint32_t blocks = 0;
    int state = 0;                      // 0 free, 1 filled, 2 processed
    bool last = false;
};

struct Pipe {
    std::mutex mu;
    std::condition_variable cv;
    Slot slot[2];
    bool failed = false;
    std::string err;
    void fail(const std::string &m)
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!failed) { failed = true; err = m; }
        cv.notify_all();
    }
    // waits until slot k reaches `state` (false: the pipeline failed)
    bool wait(int k, int state)
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return failed || slot[k].state == state; });
        return !failed;
    }
    void set(int k, int state)
    {
        std::lock_guard<std::mutex> lk(mu);
        slot[k].state = state;
        cv.notify_all();
    }
};

// fill(slot) -> false at end of input;  work(slot);  drain(slot)
template <class Fill, class Work, class Drain>
bool run_pipeline(Pipe &p, Fill fill, Work work, Drain drain)
{
    std::thread reader([&] {
        for (int k = 0;; k ^= 1) {
            if (!p.wait(k, 0)) return;
            Slot &s = p.slot[k];
            s.last = !fill(s);
            p.set(k, 1);
            if (s.last) return;
        }
    });
    std::thread writer([&] {
        for (int k = 0;; k ^= 1) {
            if (!p.wait(k, 2)) return;
            Slot &s = p.slot[k];
            const bool last = s.last;
            if (!drain(s)) { p.fail("write failed"); return; }
            p.set(k, 0);
            if (last) return;
        }
    });
    for (int k = 0;; k ^= 1) {
        if (!p.wait(k, 1)) break;
        Slot &s = p.slot[k];
        const bool last = s.last;
        if (!work(s)) { p.fail(std::string("fcx: ") + fcx_last_error()); break; }
        p.set(k, 2);
        if (last) break;
    }
    reader.join();
    writer.join();
    return !p.failed;
}

} // namespace

int main(int argc, char **argv)
{
    if (argc != 3 && argc != 4 && argc != 5 && argc != 7 && argc != 9) {           // mc:3992 (+ "-g n")
        fprintf(stderr, "usage: %s -i <input> [-o <output>] [-c lz77|lz78|fcx9] [-g <gpus>]\n", argv[0]);
        return 255;
    }
    static struct option longOpts[] = {{"file_in", required_argument, nullptr, 'i'},
                                       {"file_out", optional_argument, nullptr, 'o'},
                                       {"compress", required_argument, nullptr, 'c'},
                                       {"gpus", required_argument, nullptr, 'g'},
                                       {nullptr, 0, nullptr, 0}};
    std::string in, out = "./out", algo;                               // mc:4051-4053: default output name
    bool compress = false;
    int gpus = 1, opt;
    while ((opt = getopt_long(argc, argv, "i:o:c:g:", longOpts, nullptr)) != -1) {
        if (opt == 'i' && optarg) in = optarg;
        else if (opt == 'o' && optarg) out = optarg;
        else if (opt == 'c' && optarg) { algo = optarg; compress = true; }
        else if (opt == 'g' && optarg) gpus = atoi(optarg);
    }
    // "-c fcx9": the opt-in format of include/fcx.h (64-bit header fields, named single-symbol sections) -- the one
    // spelling that does not mean what it means to the reference (there, anything but lz77 is lz78)
    const bool x9 = compress && algo == "fcx9";
    const bool lz78 = compress && !x9 && strncmp(algo.c_str(), "lz77", 4) != 0;     // mc:4040-4043: anything else means lz78
    const int nd = fcx_device_count();
    if (nd == 0) { fprintf(stderr, "fcx: no CUDA device (this path has no CPU fallback)\n"); return 255; }
    gpus = gpus < 1 ? 1 : gpus > nd ? nd : gpus;
    const uint32_t mask = gpus >= 32 ? 0xFFFFFFFFu : ((1u << gpus) - 1u);
    FILE *fin = fopen(in.c_str(), "rb");
    if (!fin) { fprintf(stderr, "open: %s Fail!!\n", in.c_str()); return 255; }       // mc:4058-4062
    FILE *fout = fopen(out.c_str(), "wb");
    if (!fout) { fprintf(stderr, "open: %s Fail!!\n", out.c_str()); fclose(fin); return 255; }
    // FCX8: whole file in one call (compress: -c lz78; decompress: magic byte '8', mc:4153-4164)
    bool whole = lz78;
    if (!compress) {
        uint8_t magic[4] = {0, 0, 0, 0};
        const size_t k = fread(magic, 1, 4, fin);
        whole = k == 4 && memcmp(magic, "FCX8", 4) == 0;
        fseek(fin, 0, SEEK_SET);
    }
    if (whole) {
        fseek(fin, 0, SEEK_END);
        const long sz = ftell(fin);
        fseek(fin, 0, SEEK_SET);
        const uint64_t n = sz > 0 ? (uint64_t)sz : 0;
        uint16_t blocks = 0;
        uint8_t *src = static_cast<uint8_t *>(fcx_host_alloc(n + 16));
        if (!src || fread(src, 1, n, fin) != n) { fprintf(stderr, "read: %s Fail!!\n", in.c_str()); return 255; }
        if (!compress && n >= FCX_HEADER_BYTES) memcpy(&blocks, src + 8, 2);
        const uint64_t cap = compress ? fcx_compress_bound_lz78(n) : (uint64_t)(blocks ? blocks : 1) * (FCX_BLOCK_BYTES + 8);
        uint8_t *dst = static_cast<uint8_t *>(fcx_host_alloc(cap));
        uint64_t produced = 0;
        const int rc = !dst ? FCX_ERR_CUDA
                            : compress ? fcx_compress_lz78(src, n, dst, cap, &produced, mask) : fcx_decompress_lz77(src, n, dst, cap, &produced, mask);
        bool ok = rc == FCX_OK && (produced == 0 || fwrite(dst, 1, produced, fout) == produced);
        if (rc != FCX_OK) fprintf(stderr, "fcx error %d: %s\n", rc, fcx_last_error());
        fcx_host_free(src); fcx_host_free(dst);
        fclose(fin);
        if (fclose(fout) != 0) ok = false;
        return ok ? 0 : 255;
    }

    Pipe p;
    const uint64_t inCap = compress ? kPieceBytes : fcx_compress_bound(kPieceBytes);
    const uint64_t outCap = compress ? fcx_compress_bound(kPieceBytes) : kPieceBlocks * (uint64_t)(FCX_BLOCK_BYTES + 8);
    for (Slot &s : p.slot) {
        s.in = static_cast<uint8_t *>(fcx_host_alloc(inCap));
        s.out = static_cast<uint8_t *>(fcx_host_alloc(outCap));
        if (!s.in || !s.out) { fprintf(stderr, "fcx: %s\n", fcx_last_error()); return 255; }
    }
    bool ok = true;
    uint64_t totalIn = 0, totalBlocks = 0;
    auto drain = [&](Slot &s) { return s.outBytes == 0 || fwrite(s.out, 1, s.outBytes, fout) == s.outBytes; };
    if (compress) {
        uint8_t head[FCX_X9_HEADER_BYTES] = {'F', 'C', 'X', (uint8_t)(x9 ? '9' : '7'), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const size_t headLen = x9 ? FCX_X9_HEADER_BYTES : FCX_HEADER_BYTES;
        ok = fwrite(head, 1, headLen, fout) == headLen;                          // placeholder, mc:4078-4086
        if (ok) ok = run_pipeline(p,
            [&](Slot &s) {                                                       // mc:4090: fread up to one piece
                s.inBytes = fread(s.in, 1, kPieceBytes, fin);
                totalIn += s.inBytes;
                totalBlocks += (s.inBytes + FCX_BLOCK_BYTES - 1) / FCX_BLOCK_BYTES;
                return s.inBytes == kPieceBytes;
            },
            [&](Slot &s) {
                s.outBytes = 0;
                if (s.inBytes == 0) return true;
                return (x9 ? fcx_compress_records_lz77_x9(s.in, s.inBytes, s.out, outCap, &s.outBytes, mask)
                           : fcx_compress_records_lz77(s.in, s.inBytes, s.out, outCap, &s.outBytes, mask)) == FCX_OK;
            },
            drain);
        if (ok) {                                                                // mc:4128-4129: the real header
            const uint32_t tb = (uint32_t)totalIn;                               // uInt32 cmp_before_bytes wraps (mc:105)
            const uint16_t bn = (uint16_t)totalBlocks;                           // uInt16 block_num wraps (mc:106)
            if (x9) { const uint32_t bn32 = (uint32_t)totalBlocks; memcpy(head + 4, &totalIn, 8); memcpy(head + 12, &bn32, 4); }
            else { memcpy(head + 4, &tb, 4); memcpy(head + 8, &bn, 2); }
            ok = fseek(fout, 0, SEEK_SET) == 0 && fwrite(head, 1, headLen, fout) == headLen;
        }
    } else {
        uint8_t head[FCX_X9_HEADER_BYTES];
        if (fread(head, 1, FCX_HEADER_BYTES, fin) != FCX_HEADER_BYTES) { fprintf(stderr, "Read file head infomation error!!!\n"); return 255; }   // mc:4139-4144
        if (memcmp(head, "FCX", 3) != 0 || (head[3] != '7' && head[3] != '9')) { fprintf(stderr, "not an FCX7 / FCX9 file\n"); return 255; }     // mc:4146-4152
        const bool in9 = head[3] == '9';
        uint32_t left;
        if (in9) {
            if (fread(head + FCX_HEADER_BYTES, 1, 6, fin) != 6) { fprintf(stderr, "Read file head infomation error!!!\n"); return 255; }
            memcpy(&left, head + 12, 4);
        } else { uint16_t blocks; memcpy(&blocks, head + 8, 2); left = blocks; }
        ok = run_pipeline(p,
            [&](Slot &s) {                                                       // mc:4162-4174: whole records, up to one piece
                s.inBytes = 0; s.blocks = 0;
                while (left && s.blocks < kPieceBlocks) {
                    uint32_t sz;
                    if (fread(&sz, 1, 4, fin) != 4) { p.fail("truncated record header"); return false; }
                    if (s.inBytes + 4 + (uint64_t)sz > inCap) {
                        if (s.blocks == 0) { p.fail("record larger than any valid block payload"); return false; }
                        fseek(fin, -4, SEEK_CUR);                                // does not fit this piece: next one
                        break;
                    }
                    memcpy(s.in + s.inBytes, &sz, 4);
                    if (fread(s.in + s.inBytes + 4, 1, sz, fin) != sz) { p.fail("truncated record"); return false; }
                    s.inBytes += 4 + (uint64_t)sz;
                    s.blocks++; left--;
                }
                return left != 0;
            },
            [&](Slot &s) {
                s.outBytes = 0;
                if (s.blocks == 0) return true;
                return (in9 ? fcx_decompress_records_lz77_x9(s.in, s.inBytes, s.blocks, s.out, outCap, &s.outBytes, mask)
                            : fcx_decompress_records_lz77(s.in, s.inBytes, s.blocks, s.out, outCap, &s.outBytes, mask)) == FCX_OK;
            },
            drain);
    }
    if (!ok) fprintf(stderr, "%s\n", p.failed ? p.err.c_str() : "write failed");
    for (Slot &s : p.slot) { fcx_host_free(s.in); fcx_host_free(s.out); }
    fclose(fin);
    if (fclose(fout) != 0) ok = false;
    return ok ? 0 : 255;
}
