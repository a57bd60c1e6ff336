// Input side of vcpenc_transcode: a source of decoded pictures in one of K1's input layouts.
#ifndef VCP_FRONTEND_H
#define VCP_FRONTEND_H

#include <stddef.h>
#include <stdint.h>
#include <unistd.h>

#include <memory>
#include <mutex>

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
    // container inputs: the audio that goes into the output beside the video (access units accumulate in
    // audio.data while read() demuxes; the caller drains them) and an estimate of the number of pictures (0: unknown)
    AudioTrack audio;
    std::mutex audio_mu;   // guards audio.data / audio.sizes (the reader thread appends, the muxing thread drains)
    long est_frames = 0;
    long exact_frames = 0;   // raw / y4m files: the number of pictures, known up front (0: unknown) -- lets the caller upload while it reads
    // end of input reached: flush the audio encoder (call once, after the last read())
    virtual int finish_audio(char*, size_t) { return 0; }
};

// container files (.mp4 .mkv .avi .mov .webm ...): demux + decode through libavformat / libavcodec
// loaded at run time (frontend_lav.cpp)
// audio_bitrate: `-b:a` of the preset (bits per second) for inputs whose audio has to be re-encoded to AAC
int open_container_source(const char* path, bool drop_audio, int audio_bitrate, std::unique_ptr<FrameSource>* out, char* err, size_t errlen);

}  // namespace vcp

#endif
