// Small host-side helpers shared by args / mux / transcode / verify.
#ifndef VCP_HOST_UTIL_H
#define VCP_HOST_UTIL_H

#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/vcpenc.h"

namespace vcp {

void set_err(char* err, size_t errlen, const char* fmt, ...);

struct NalRef { const uint8_t* p; size_t n; };
// split an Annex-B buffer into NAL units (start codes removed)
std::vector<NalRef> split_annexb(const uint8_t* d, size_t n);

// MP4 writer (mux_mp4.cpp)
struct Mp4Sample { uint64_t offset; uint32_t size; bool sync; };
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
int write_mp4(const vcpenc_params& p, const ParamSets& ps,
              const std::vector<Mp4Sample>& samples, const uint8_t* mdat, uint64_t mdat_len,
              const char* path, char* err, size_t errlen);

}  // namespace vcp

#endif
