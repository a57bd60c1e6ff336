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
int write_mp4(const vcpenc_params& p, const std::vector<uint8_t>& sps, const std::vector<uint8_t>& pps,
              const std::vector<Mp4Sample>& samples, const uint8_t* mdat, uint64_t mdat_len,
              const char* path, char* err, size_t errlen);

}  // namespace vcp

#endif
