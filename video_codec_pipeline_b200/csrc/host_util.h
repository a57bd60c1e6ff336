// Small host-side helpers shared by args / mux / transcode / verify.
#ifndef VCP_HOST_UTIL_H
#define VCP_HOST_UTIL_H

#include <stddef.h>
#include <stdint.h>

#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/vcpenc.h"

namespace vcp {

void set_err(char* err, size_t errlen, const char* fmt, ...);

struct NalRef { const uint8_t* p; size_t n; };
// split an Annex-B buffer into NAL units (start codes removed)
std::vector<NalRef> split_annexb(const uint8_t* d, size_t n);

// parameter sets of the stream (NAL payloads incl. header); take() keeps the first of each kind and says
// whether the NAL was one (parameter sets live in avcC / hvcC, not in the samples)
struct ParamSets {
    std::vector<uint8_t> vps, sps, pps;   // vps: HEVC only
    bool take(int codec, const NalRef& nal) {
        std::vector<uint8_t>* dst = nullptr;
        if (codec == VCPENC_CODEC_HEVC) { const int t = (nal.p[0] >> 1) & 63; dst = t == 32 ? &vps : t == 33 ? &sps : t == 34 ? &pps : nullptr; }
        else { const int t = nal.p[0] & 31; dst = t == 7 ? &sps : t == 8 ? &pps : nullptr; }
        if (!dst) return false;
        if (dst->empty()) dst->assign(nal.p, nal.p + nal.n);
        return true;
    }
};

// The AAC track that goes beside the video (every encode preset carries `-c:a aac -b:a Nk`,
// /root/reference/internal/config/config.go:45-50).  Filled by the container front end: frames are raw AAC
// access units (no ADTS), either copied from an AAC-LC input or produced by libavcodec's `aac` encoder.
struct AudioTrack {
    bool present = false;
    int sample_rate = 0, channels = 0;
    int frame_samples = 1024;        // samples per access unit
    int priming = 0;                 // encoder delay at the start of the track, in samples (edit list)
    int bitrate = 0;
    bool copied = false;             // stream copy of the input's AAC (else decoded and re-encoded)
    std::vector<uint8_t> asc;        // AudioSpecificConfig for esds
    std::vector<uint8_t> data;       // access units not yet handed to the muxer, back to back
    std::vector<uint32_t> sizes;
    uint64_t frames_total = 0;
};

// MP4 writer (mux_mp4.cpp): payload streamed to the file, moov at the end (or into the space reserved in front)
struct Mp4Chunk { uint64_t offset; uint32_t nsamples; };
class Mp4Writer {
public:
    ~Mp4Writer();
    // expect_*: samples the caller expects (0 = unknown), sizes the moov reserve of a faststart file
    int open(const char* path, const vcpenc_params& p, uint64_t expect_vsamples, uint64_t expect_asamples, char* err, size_t errlen);
    int video_access_unit(const uint8_t* annexb, size_t len, bool sync);   // Annex-B in, length-prefixed sample out
    int audio_frame(const uint8_t* d, size_t n);
    void end_chunk();                                                       // the next sample of either track starts a new chunk
    int finish(const AudioTrack* audio, char* err, size_t errlen);          // writes moov, closes the file
    void abandon();                                                         // close + remove
    uint64_t video_samples() const { return vsize_.size(); }
private:
    std::vector<uint8_t> build_moov(uint64_t shift) const;
    bool put(int track, const uint8_t* d, size_t n);
    FILE* f_ = nullptr;
    std::string path_;
    vcpenc_params p_{};
    ParamSets ps_;
    AudioTrack audio_;
    std::vector<uint32_t> vsize_, vsync_, asize_;
    std::vector<Mp4Chunk> vchunks_, achunks_;
    std::vector<uint8_t> scratch_;
    std::vector<char> iobuf_;
    uint64_t pos_ = 0, moov_at_ = 0, reserve_ = 0, mdat_at_ = 0;
    int cur_track_ = -1;
    bool failed_ = false;
};

}  // namespace vcp

#endif
