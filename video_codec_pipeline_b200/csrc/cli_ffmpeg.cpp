// vcp-ffmpeg — accepts exactly the argv the reference consumer builds
//   ffmpeg -hide_banner -loglevel warning -y -i INPUT <tokens...> OUTPUT
// (/root/reference/cmd/consumer.go:376-382) and runs it on the B200 encoder.  Installed as
// `ffmpeg` earlier on $PATH it makes a stock, unmodified `vcp` binary use libvcpenc.
// Exit status: 0 on success, the vcpenc error class otherwise (the consumer only tests != 0).
// SIGTERM/SIGINT set the cancel flag (exec.CommandContext kills the child on ctx cancel).
#include <csignal>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vcpenc.h"

static volatile int g_cancel = 0;
static void on_signal(int) { g_cancel = 1; }

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
    if (rc) fprintf(stderr, "vcp-ffmpeg: error class %d: %s\n", rc, err);
    return rc;
}
