// HEVC luma sample interpolation at quarter-sample positions (8.5.3.3.3.1, 8-bit: shift1 = 0, shift2 = 6, then the
// default weighted prediction (x + 32) >> 6) for one 16x16 block, by one warp, out of a window of integer samples
// in shared memory.  Oracle: hevc_luma_pred (oracle/hevc_oracle.inc.c).
//
// Every position goes through the separable path: rows first (8-tap, 16-bit intermediates T), then columns
// ((sum >> 6 + 32) >> 6).  With the fraction-0 "filter" {0,0,0,64,0,0,0,0} on one axis this equals the
// one-dimensional formulas of the standard exactly (64 t >> 6 = t), so there are no special cases.
//
// The row pass is shared by every candidate with the same horizontal position: it writes W[R][x] = (T[R][x], T[R+1][x])
// as two 16-bit halves for all window rows R, so that the column pass of ANY start row is four dp2a per sample
// (rows s..s+7 = the words W[s], W[s+2], W[s+4], W[s+6]) whatever the parity of s.
#ifndef VCP_HEVC_QPEL_CUH
#define VCP_HEVC_QPEL_CUH

#include <cuda_runtime.h>
#include <stdint.h>

constexpr int HQ_WIN = 24;        // window rows / columns: samples -4 .. 19 around the block at its full-sample vector
constexpr int HQ_WPITCH = 20;     // words per row of W (16 used)
constexpr int HQ_WROWS = HQ_WIN - 1;

__host__ __device__ constexpr uint32_t hq_pack4(int a, int b, int c, int d) {
    return (uint32_t)(a & 255) | ((uint32_t)(b & 255) << 8) | ((uint32_t)(c & 255) << 16) | ((uint32_t)(d & 255) << 24);
}
// Table 8-11, taps 0..3 and 4..7 as four signed bytes each
__device__ __forceinline__ void hq_filter(int f, uint32_t& lo, uint32_t& hi) {
    lo = f == 0 ? hq_pack4(0, 0, 0, 64) : f == 1 ? hq_pack4(-1, 4, -10, 58) : f == 2 ? hq_pack4(-1, 4, -11, 40) : hq_pack4(0, 1, -5, 17);
    hi = f == 0 ? hq_pack4(0, 0, 0, 0) : f == 1 ? hq_pack4(17, -5, 1, 0) : f == 2 ? hq_pack4(40, -11, 4, -1) : hq_pack4(58, -10, 4, -1);
}
__device__ __forceinline__ int hq_dp4a_us(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// eight row-filtered samples: out[x] = sum_i f[i] * p[x + i], p = the 15 bytes starting at byte `o` of the row `r`
__device__ __forceinline__ void hq_row8(const uint32_t* __restrict__ r, int o, uint32_t flo, uint32_t fhi, int out[8]) {
    const uint32_t* q = r + (o >> 2);
    const uint32_t sh = (uint32_t)(o & 3) * 8u;
    const uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
    uint32_t w[4] = {__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh), __funnelshift_r(a3, a4, sh)};
    uint32_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = (k & 3) ? __funnelshift_r(w[k >> 2], w[(k >> 2) + 1], 8u * (k & 3)) : w[k >> 2];
#pragma unroll
    for (int x = 0; x < 8; x++) out[x] = hq_dp4a_us(s[x + 4], fhi, hq_dp4a_us(s[x], flo, 0));
}
// Row pass.  win: first word of window row 0, pitchw words per row; byte0: byte inside a row of the first tap of output
// column 0 (window column ixo + 1 plus the window's misalignment); fx: horizontal fraction 0..3.
__device__ __forceinline__ void hq_hpass(const uint32_t* __restrict__ win, int pitchw, int byte0, int fx, uint32_t* __restrict__ W, int lane) {
    uint32_t flo, fhi;
    hq_filter(fx, flo, fhi);
    const int half = lane & 1;
#pragma unroll 1
    for (int R = lane >> 1; R < HQ_WROWS; R += 16) {
        int t0[8], t1[8];
        hq_row8(win + R * pitchw, byte0 + 8 * half, flo, fhi, t0);
        hq_row8(win + (R + 1) * pitchw, byte0 + 8 * half, flo, fhi, t1);
        uint32_t p[8];
#pragma unroll
        for (int x = 0; x < 8; x++) p[x] = ((uint32_t)t0[x] & 0xffffu) | ((uint32_t)t1[x] << 16);
        uint4* dst = reinterpret_cast<uint4*>(W + R * HQ_WPITCH + 8 * half);
        dst[0] = make_uint4(p[0], p[1], p[2], p[3]);
        dst[1] = make_uint4(p[4], p[5], p[6], p[7]);
    }
}
// Column pass: the eight predicted samples x = 8 half .. 8 half + 7 of the block row whose first tap is W row s; fy 0..3
__device__ __forceinline__ uint2 hq_vpass(const uint32_t* __restrict__ W, int s, int half, int fy) {
    uint32_t glo, ghi;
    hq_filter(fy, glo, ghi);
    int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint4* src = reinterpret_cast<const uint4*>(W + (s + 2 * j) * HQ_WPITCH + 8 * half);
        const uint4 u = src[0], v = src[1];
        const uint32_t w[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
        const uint32_t gsel = j < 2 ? glo : ghi;
#pragma unroll
        for (int x = 0; x < 8; x++) acc[x] = (j & 1) ? __dp2a_hi((int)w[x], (int)gsel, acc[x]) : __dp2a_lo((int)w[x], (int)gsel, acc[x]);
    }
    uint32_t o[2] = {0u, 0u};
#pragma unroll
    for (int x = 0; x < 8; x++) {
        int v = ((acc[x] >> 6) + 32) >> 6;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        o[x >> 2] |= (uint32_t)v << (8 * (x & 3));
    }
    return make_uint2(o[0], o[1]);
}

#endif  // VCP_HEVC_QPEL_CUH
