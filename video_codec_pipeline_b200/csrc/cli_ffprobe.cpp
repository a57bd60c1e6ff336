// vcp-ffprobe — accepts the argv of the reference's --verify step
//   ffprobe -v error -select_streams v:0 -show_entries stream=codec_type -of csv=p=0 PATH
// (/root/reference/cmd/consumer.go:409-410) and prints "video" when PATH holds a video
// stream; the consumer passes iff the exit status is 0 and stdout contains "video" (:411-418).
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/vcpenc.h"

int main(int argc, char** argv) {
    std::string path;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-v" || a == "-loglevel" || a == "-select_streams" || a == "-show_entries" || a == "-of" || a == "-print_format") { i++; continue; }
        if (a == "-hide_banner") continue;
        path = a;
    }
    if (path.empty()) { fprintf(stderr, "usage: vcp-ffprobe [options] PATH\n"); return 1; }
    char err[512] = {0};
    const int rc = vcpenc_verify(path.c_str(), err, sizeof err);
    if (rc) { fprintf(stderr, "%s: %s\n", path.c_str(), err); return 1; }
    printf("video\n");
    return 0;
}
