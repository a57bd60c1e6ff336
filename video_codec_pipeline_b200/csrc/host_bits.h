// Host-side bit writer + parameter-set (SPS/PPS) writers used by the encoder core and the
// MP4 muxer.  Mirrors the syntax of H.264 7.3.2.1 / 7.3.2.2; the slice layer is written on
// the device (k5_cavlc.cu).
#ifndef VCP_HOST_BITS_H
#define VCP_HOST_BITS_H

#include <stdint.h>
#include <vector>

#include "../../include/vcpenc.h"

namespace vcp {

class BitWriter {
public:
    void put(int n, uint32_t v) {
        for (int i = n - 1; i >= 0; i--) {
            cur_ = (uint8_t)((cur_ << 1) | ((v >> i) & 1));
            if (++nb_ == 8) { buf_.push_back(cur_); cur_ = 0; nb_ = 0; }
        }
    }
    void put32(uint32_t v) { put(16, v >> 16); put(16, v & 0xffff); }
    void ue(uint32_t k) {
        uint32_t x = k + 1; int n = 0;
        while ((x >> n) > 1) n++;
        put(n, 0); put(n + 1, x);
    }
    void se(int v) { ue(v <= 0 ? (uint32_t)(-2 * v) : (uint32_t)(2 * v - 1)); }
    void trailing() { put(1, 1); if (nb_) put(8 - nb_, 0); }
    const std::vector<uint8_t>& bytes() const { return buf_; }
private:
    std::vector<uint8_t> buf_;
    uint8_t cur_ = 0;
    int nb_ = 0;
};

// RBSP -> NAL payload with emulation prevention (no start code)
inline std::vector<uint8_t> nal_escape(int ref_idc, int type, const std::vector<uint8_t>& rbsp) {
    std::vector<uint8_t> o;
    o.push_back((uint8_t)((ref_idc << 5) | type));
    int zeros = 0;
    for (uint8_t b : rbsp) {
        if (zeros >= 2 && b <= 3) { o.push_back(3); zeros = 0; }
        o.push_back(b);
        zeros = b == 0 ? zeros + 1 : 0;
    }
    return o;
}

// HEVC: two-byte nal_unit_header (layer 0, temporal id 0), same emulation prevention
inline std::vector<uint8_t> hevc_nal_escape(int type, const std::vector<uint8_t>& rbsp) {
    std::vector<uint8_t> o;
    o.push_back((uint8_t)(type << 1));
    o.push_back(1);
    int zeros = 0;
    for (uint8_t b : rbsp) {
        if (zeros >= 2 && b <= 3) { o.push_back(3); zeros = 0; }
        o.push_back(b);
        zeros = b == 0 ? zeros + 1 : 0;
    }
    return o;
}

int level_idc_for(int mbw, int mbh, int fps_num, int fps_den);
std::vector<uint8_t> make_sps_nal(const vcpenc_params& p);  // NAL payload incl. header byte
std::vector<uint8_t> make_pps_nal(const vcpenc_params& p);
// HEVC parameter sets for the stream structure of k6_hevc.cu (NAL payload incl. the two header bytes)
int hevc_level_idc_for(int cw, int ch, int fps_num, int fps_den);
std::vector<uint8_t> make_hevc_vps_nal(const vcpenc_params& p);
std::vector<uint8_t> make_hevc_sps_nal(const vcpenc_params& p);
std::vector<uint8_t> make_hevc_pps_nal(const vcpenc_params& p);

}  // namespace vcp

#endif
