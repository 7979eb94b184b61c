// fcx_cli -- flag-compatible stand-in for the reference CLI (SURVEY 8f "next" row 2):
//     ./fcx_cli -i <in> [-o <out>] [-c lz77]        (mc:3992-4053: no -c => decompress, default out "./out")
// The whole file goes through fcx_compress_lz77 / fcx_decompress_lz77 in one batched device pass
// instead of the reference's block-at-a-time loop (mc:4090-4122); the bytes written are identical.
// Unlike the reference it prints nothing but errors, and "-c <anything else>" (lz78) is refused.
#include "../../include/fcx.h"

#include <getopt.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static bool read_file(const char *path, std::vector<uint8_t> &buf)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    bool ok = n <= 0 || fread(buf.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

int main(int argc, char **argv)
{
    if (argc != 3 && argc != 4 && argc != 5 && argc != 7) {           // mc:3992
        fprintf(stderr, "usage: %s -i <input> [-o <output>] [-c lz77]\n", argv[0]);
        return 255;
    }
    static struct option longOpts[] = {{"file_in", required_argument, nullptr, 'i'},
                                       {"file_out", optional_argument, nullptr, 'o'},
                                       {"compress", required_argument, nullptr, 'c'},
                                       {nullptr, 0, nullptr, 0}};
    std::string in, out = "./out", algo;                               // mc:4051-4053: default output name
    bool compress = false;
    int opt;
    while ((opt = getopt_long(argc, argv, "i:o:c:", longOpts, nullptr)) != -1) {
        if (opt == 'i' && optarg) in = optarg;
        else if (opt == 'o' && optarg) out = optarg;
        else if (opt == 'c' && optarg) { algo = optarg; compress = true; }
    }
    std::vector<uint8_t> src;
    if (!read_file(in.c_str(), src)) { fprintf(stderr, "open: %s Fail!!\n", in.c_str()); return 255; }
    std::vector<uint8_t> dst;
    uint64_t produced = 0;
    int rc;
    if (compress) {
        if (strncmp(algo.c_str(), "lz77", 4) != 0) {                   // mc:4040-4043: anything else means lz78
            fprintf(stderr, "only -c lz77 is on this path (lz78 containers are out of scope)\n");
            return 255;
        }
        dst.resize(fcx_compress_bound(src.size()));
        rc = fcx_compress_lz77(src.data(), src.size(), dst.data(), dst.size(), &produced, 0);
    } else {
        if (src.size() < FCX_HEADER_BYTES) { fprintf(stderr, "Read file head infomation error!!!\n"); return 255; }
        uint16_t blocks;
        memcpy(&blocks, src.data() + 8, 2);
        dst.resize((size_t)(blocks ? blocks : 1) * (FCX_BLOCK_BYTES + 8));
        rc = fcx_decompress_lz77(src.data(), src.size(), dst.data(), dst.size(), &produced, 0);
    }
    if (rc != FCX_OK) { fprintf(stderr, "fcx error %d: %s\n", rc, fcx_last_error()); return 255; }
    FILE *f = fopen(out.c_str(), "wb");
    if (!f) { fprintf(stderr, "open: %s Fail!!\n", out.c_str()); return 255; }
    if (produced) fwrite(dst.data(), 1, produced, f);
    fclose(f);
    return 0;
}
