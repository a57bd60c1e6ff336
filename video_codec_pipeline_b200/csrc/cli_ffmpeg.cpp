// vcp-ffmpeg — accepts exactly the argv the reference consumer builds
//   ffmpeg -hide_banner -loglevel warning -y -i INPUT <tokens...> OUTPUT
// (/root/reference/cmd/consumer.go:376-382) and runs it on the B200 encoder.  Installed as
// `ffmpeg` earlier on $PATH it makes a stock, unmodified `vcp` binary use libvcpenc.
// Exit status: 0 on success, the vcpenc error class otherwise (the consumer only tests != 0).
// exec.CommandContext SIGKILLs the child on cancel / timeout (nothing to handle: the output file is only
// complete when we return 0, and the consumer removes it on failure, cmd/consumer.go:264); an operator's
// SIGTERM / SIGINT sets the cancel flag so the partial output is removed before exiting.
// Tasks that are not a B200 video encode -- `-c copy`, `-vn` (VCPENC_E_NOTENCODE), audio the loaded FFmpeg
// libraries cannot carry (VCPENC_E_AUDIO), VCPENC_E_UNSUPPORTED -- are handed to the stock ffmpeg with the
// ORIGINAL argv: $VCP_STOCK_FFMPEG, else the next `ffmpeg` on $PATH that is not this program.
#include <climits>
#include <csignal>
#include <cstdlib>
#include <unistd.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vcpenc.h"

static volatile int g_cancel = 0;
static void on_signal(int) { g_cancel = 1; }

// the stock ffmpeg to hand non-encode tasks to, or "" if there is none
static std::string stock_ffmpeg() {
    if (const char* e = getenv("VCP_STOCK_FFMPEG")) return e;
    char self[PATH_MAX] = {0};
    if (!realpath("/proc/self/exe", self)) self[0] = 0;
    const char* path = getenv("PATH");
    if (!path) return "";
    std::string p = path;
    size_t a = 0;
    while (a <= p.size()) {
        size_t b = p.find(':', a);
        if (b == std::string::npos) b = p.size();
        const std::string cand = (b > a ? p.substr(a, b - a) : std::string(".")) + "/ffmpeg";
        char real[PATH_MAX] = {0};
        if (access(cand.c_str(), X_OK) == 0 && realpath(cand.c_str(), real) && strcmp(real, self) != 0) return cand;
        a = b + 1;
    }
    return "";
}

int main(int argc, char** argv) {
    signal(SIGTERM, on_signal);
    signal(SIGINT, on_signal);
    std::string input, output;
    std::vector<const char*> toks;
    bool seen_input = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (!seen_input) {
            // global / input options in front of -i
            if (a == "-hide_banner" || a == "-y" || a == "-nostdin") continue;
            if (a == "-loglevel" || a == "-v") { i++; continue; }
            if (a == "-i") {
                if (i + 1 >= argc) { fprintf(stderr, "vcp-ffmpeg: -i needs a path\n"); return VCPENC_E_ARGS; }
                input = argv[++i]; seen_input = true; continue;
            }
            if (a == "-version") { printf("%s\n", vcpenc_version()); return 0; }
            if (a == "-encoders") { printf(" V..... h264_nvenc           B200 CUDA H.264 encoder (libvcpenc)\n V..... libx264              B200 CUDA H.264 encoder (libvcpenc)\n V..... hevc_nvenc           B200 CUDA HEVC encoder (libvcpenc)\n V..... libx265              B200 CUDA HEVC encoder (libvcpenc)\n"); return 0; }
            // options that take a value and precede -i (-s, -r, -f, -pix_fmt for raw input)
            toks.push_back(argv[i]);
            continue;
        }
        if (i == argc - 1) { output = a; break; }
        toks.push_back(argv[i]);
    }
    if (input.empty() || output.empty()) {
        fprintf(stderr, "usage: vcp-ffmpeg [-hide_banner] [-loglevel L] [-y] -i INPUT [options] OUTPUT\n");
        return VCPENC_E_ARGS;
    }
    char err[1024] = {0};
    const int rc = vcpenc_transcode(input.c_str(), output.c_str(), (int)toks.size(), toks.data(), 0, &g_cancel, err, sizeof err);
    if (rc == VCPENC_E_NOTENCODE || rc == VCPENC_E_AUDIO || rc == VCPENC_E_UNSUPPORTED) {
        const std::string stock = stock_ffmpeg();
        if (!stock.empty()) {
            fprintf(stderr, "vcp-ffmpeg: %s -- handing the task to %s\n", err, stock.c_str());
            argv[0] = const_cast<char*>(stock.c_str());
            execv(stock.c_str(), argv);
        }
        fprintf(stderr, "vcp-ffmpeg: no stock ffmpeg to hand the task to (set VCP_STOCK_FFMPEG)\n");
    }
    if (rc) fprintf(stderr, "vcp-ffmpeg: error class %d: %s\n", rc, err);
    return rc;
}
