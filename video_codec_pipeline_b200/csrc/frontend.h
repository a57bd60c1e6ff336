// Input side of vcpenc_transcode: a source of decoded pictures in one of K1's input layouts.
#ifndef VCP_FRONTEND_H
#define VCP_FRONTEND_H

#include <stddef.h>
#include <stdint.h>
#include <unistd.h>

#include <memory>

#include "host_util.h"
#include "vcp_algo.h"

namespace vcp {

struct FrameSource {
    virtual ~FrameSource() {}
    int width = 0, height = 0, fps_num = 0, fps_den = 1;
    int fmt = VCPENC_FMT_YUV420P;
    size_t fbytes() const { return (size_t)vcp_in_frame_bytes(fmt, width, height); }
    // read up to `max` frames (tight, `fmt`) into dst; returns frames read, <0 = -(error class)
    virtual int read(uint8_t* dst, int max, char* err, size_t errlen) = 0;
};

// container files (.mp4 .mkv .avi .mov .webm ...): demux + decode through libavformat / libavcodec
// loaded at run time (frontend_lav.cpp)
int open_container_source(const char* path, bool drop_audio, std::unique_ptr<FrameSource>* out, char* err, size_t errlen);

}  // namespace vcp

#endif
