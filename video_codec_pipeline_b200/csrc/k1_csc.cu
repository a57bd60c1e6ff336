// K1 — colour-space / layout convert into the encoder's padded yuv420p planes, plus the
// half-resolution luma plane used by the motion-search pre-pass.
//
// Replaces the auto-inserted `scale`/`format` filter inside the ffmpeg child that the
// reference spawns (/root/reference/cmd/consumer.go:376-382).  Pure streaming, HBM-bound:
// every thread moves 16-byte vectors; borders are produced by clamping the source
// coordinate, so no second pass over the picture is needed.
//
// Algorithmic bytes per frame (yuv420p in): read 1.5*W*H + write 1.5*W*H; the
// half-res plane and the borders are extra (~ +20 %) and are counted as overhead.
#include "../../include/vcpenc.h"
#include "vcp_dev.cuh"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// load 16 output bytes of a padded row: out x in [X0, X0+16), source row `src` of width w
__device__ __forceinline__ uint4 load16_clamped(const uint8_t* __restrict__ src, int X0, int w) {
    if (X0 >= 0 && X0 + 15 < w && ((reinterpret_cast<uintptr_t>(src + X0) & 15) == 0))
        return __ldg(reinterpret_cast<const uint4*>(src + X0));
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) v |= (uint32_t)__ldg(src + clampi(X0 + 4 * k + j, 0, w - 1)) << (8 * j);
        r[k] = v;
    }
    return make_uint4(r[0], r[1], r[2], r[3]);
}

// average 2x2 -> one byte; inputs are two rows of 8 bytes (a,b as 2 words each) -> 4 bytes
__device__ __forceinline__ uint32_t half4(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t wa = i < 2 ? a0 : a1, wb = i < 2 ? b0 : b1;
        int sh = (i & 1) * 16;
        uint32_t s = ((wa >> sh) & 255) + ((wa >> (sh + 8)) & 255) + ((wb >> sh) & 255) + ((wb >> (sh + 8)) & 255);
        out |= ((s + 2) >> 2) << (8 * i);
    }
    return out;
}

// grid: x = tiles of 32 luma px over the stride, y = row pairs of the padded luma plane
// (== rows of the padded chroma planes), z = frame
__global__ void __launch_bounds__(64) k1_yuv420p_kernel(const uint8_t* __restrict__ in, size_t frame_bytes, int n0,
                                                        VcpGeom g, VcpBufs b) {
    int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile * 32 >= g.ys) return;
    int rp = blockIdx.y;
    int n = n0 + blockIdx.z;
    const uint8_t* f = in + (size_t)blockIdx.z * frame_bytes;
    int w = g.w, h = g.h, cw = (w + 1) >> 1, chh = (h + 1) >> 1;
    const uint8_t* sy = f;
    const uint8_t* su = f + (size_t)w * h;
    const uint8_t* sv = su + (size_t)cw * chh;

    // luma: rows 2rp, 2rp+1 of the padded plane
    int X0 = tile * 32 - VCP_PAD;
    uint4 r0[2], r1[2];
    {
        int y0 = clampi(2 * rp - VCP_PAD, 0, h - 1), y1 = clampi(2 * rp + 1 - VCP_PAD, 0, h - 1);
        r0[0] = load16_clamped(sy + (size_t)y0 * w, X0, w);
        r0[1] = load16_clamped(sy + (size_t)y0 * w, X0 + 16, w);
        r1[0] = load16_clamped(sy + (size_t)y1 * w, X0, w);
        r1[1] = load16_clamped(sy + (size_t)y1 * w, X0 + 16, w);
        uint8_t* dy = b.src_y + (size_t)n * g.ysize + (size_t)(2 * rp) * g.ys + tile * 32;
        reinterpret_cast<uint4*>(dy)[0] = r0[0];
        reinterpret_cast<uint4*>(dy)[1] = r0[1];
        reinterpret_cast<uint4*>(dy + g.ys)[0] = r1[0];
        reinterpret_cast<uint4*>(dy + g.ys)[1] = r1[1];
        uint4 hv;
        hv.x = half4(r0[0].x, r0[0].y, r1[0].x, r1[0].y);
        hv.y = half4(r0[0].z, r0[0].w, r1[0].z, r1[0].w);
        hv.z = half4(r0[1].x, r0[1].y, r1[1].x, r1[1].y);
        hv.w = half4(r0[1].z, r0[1].w, r1[1].z, r1[1].w);
        uint8_t* dh = b.src_h + (size_t)n * g.hsize + (size_t)rp * g.hs + tile * 16;
        *reinterpret_cast<uint4*>(dh) = hv;
    }
    // chroma: row rp of the padded planes
    {
        int yc = clampi(rp - VCP_PADC, 0, chh - 1);
        int XC = tile * 16 - VCP_PADC;
        uint4 u = load16_clamped(su + (size_t)yc * cw, XC, cw);
        uint4 v = load16_clamped(sv + (size_t)yc * cw, XC, cw);
        *reinterpret_cast<uint4*>(b.src_u + (size_t)n * g.csize + (size_t)rp * g.cs + tile * 16) = u;
        *reinterpret_cast<uint4*>(b.src_v + (size_t)n * g.csize + (size_t)rp * g.cs + tile * 16) = v;
    }
}


// ---- other input formats and scaling ---------------------------------------------------------
// Two optional front stages normalise what ffmpeg's auto-inserted swscale would: any accepted
// pixel format -> tight yuv420p of the input size, then bilinear resampling to the output size.
// Both restate vcp_algo.h (vcp_rgb_*, vcp_scale_pos, vcp_bilerp) so the oracle computes the
// same bytes.  They write a scratch picture that the layout kernel above then pads; yuv420p at
// the output size (the north-star case) skips both.

// thread = one chroma sample = one 2x2 luma quad; grid z = frame
__global__ void __launch_bounds__(128) k1_to_yuv420p_kernel(const uint8_t* __restrict__ in, size_t in_fb, int fmt,
                                                            int w, int h, uint8_t* __restrict__ out, size_t out_fb) {
    const int cw = (w + 1) >> 1, chh = (h + 1) >> 1;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cw || y >= chh) return;
    const uint8_t* f = in + (size_t)blockIdx.z * in_fb;
    uint8_t* Y = out + (size_t)blockIdx.z * out_fb;
    uint8_t* U = Y + (size_t)w * h;
    uint8_t* V = U + (size_t)cw * chh;
    const int x0 = 2 * x, y0 = 2 * y;
    const int x1 = x0 + 1 < w ? x0 + 1 : w - 1, y1 = y0 + 1 < h ? y0 + 1 : h - 1;
    const bool hx = x0 + 1 < w, hy = y0 + 1 < h;
    const size_t wh = (size_t)w * h;
    if (fmt == VCPENC_FMT_RGB24 || fmt == VCPENC_FMT_BGR24) {
        const int ro = fmt == VCPENC_FMT_RGB24 ? 0 : 2, bo = 2 - ro;
        int r = 0, gsum = 0, bsum = 0;
#pragma unroll
        for (int dy = 0; dy < 2; dy++)
#pragma unroll
            for (int dx = 0; dx < 2; dx++) {
                const int xx = dx ? x1 : x0, yy = dy ? y1 : y0;
                const uint8_t* p = f + ((size_t)yy * w + xx) * 3;
                const int R = __ldg(p + ro), G = __ldg(p + 1), B = __ldg(p + bo);
                r += R; gsum += G; bsum += B;
                if ((!dx || hx) && (!dy || hy)) Y[(size_t)yy * w + xx] = (uint8_t)vcp_rgb_y(R, G, B);
            }
        r = (r + 2) >> 2; gsum = (gsum + 2) >> 2; bsum = (bsum + 2) >> 2;
        U[(size_t)y * cw + x] = (uint8_t)vcp_rgb_u(r, gsum, bsum);
        V[(size_t)y * cw + x] = (uint8_t)vcp_rgb_v(r, gsum, bsum);
        return;
    }
    // planar / semi-planar: luma is a copy
    Y[(size_t)y0 * w + x0] = __ldg(f + (size_t)y0 * w + x0);
    if (hx) Y[(size_t)y0 * w + x1] = __ldg(f + (size_t)y0 * w + x1);
    if (hy) {
        Y[(size_t)y1 * w + x0] = __ldg(f + (size_t)y1 * w + x0);
        if (hx) Y[(size_t)y1 * w + x1] = __ldg(f + (size_t)y1 * w + x1);
    }
    if (fmt == VCPENC_FMT_NV12) {
        const uint8_t* uv = f + wh + ((size_t)y * cw + x) * 2;
        U[(size_t)y * cw + x] = __ldg(uv);
        V[(size_t)y * cw + x] = __ldg(uv + 1);
    } else if (fmt == VCPENC_FMT_YUV444P) {
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            const uint8_t* s = f + wh * (1 + pl);
            const int v = (__ldg(s + (size_t)y0 * w + x0) + __ldg(s + (size_t)y0 * w + x1) + __ldg(s + (size_t)y1 * w + x0) +
                           __ldg(s + (size_t)y1 * w + x1) + 2) >> 2;
            (pl ? V : U)[(size_t)y * cw + x] = (uint8_t)v;
        }
    } else if (fmt == VCPENC_FMT_YUV422P) {
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            const uint8_t* s = f + wh + (size_t)pl * cw * h;
            const int v = (__ldg(s + (size_t)y0 * cw + x) + __ldg(s + (size_t)y1 * cw + x) + 1) >> 1;
            (pl ? V : U)[(size_t)y * cw + x] = (uint8_t)v;
        }
    } else {  // yuv420p
        U[(size_t)y * cw + x] = __ldg(f + wh + (size_t)y * cw + x);
        V[(size_t)y * cw + x] = __ldg(f + wh + (size_t)cw * chh + (size_t)y * cw + x);
    }
}

// bilinear resample of tight yuv420p; thread = one output sample, grid y = rows of Y then U then V
__global__ void __launch_bounds__(128) k1_scale_kernel(const uint8_t* __restrict__ in, size_t in_fb, int sw, int sh,
                                                       uint8_t* __restrict__ out, size_t out_fb, int dw, int dh) {
    const int scw = (sw + 1) >> 1, sch = (sh + 1) >> 1, dcw = (dw + 1) >> 1, dch = (dh + 1) >> 1;
    int row = blockIdx.y;
    const uint8_t* s = in + (size_t)blockIdx.z * in_fb;
    uint8_t* d = out + (size_t)blockIdx.z * out_fb;
    int pw = sw, ph = sh, qw = dw, qh = dh;
    if (row >= dh) {
        row -= dh; s += (size_t)sw * sh; d += (size_t)dw * dh;
        pw = scw; ph = sch; qw = dcw; qh = dch;
        if (row >= dch) { row -= dch; s += (size_t)scw * sch; d += (size_t)dcw * dch; }
    }
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= qw) return;
    const long long py = vcp_scale_pos(row, ph, qh), px = vcp_scale_pos(x, pw, qw);
    const int yy0 = (int)(py >> 16), fy = (int)((py & 0xffff) >> 8), yy1 = yy0 + 1 < ph ? yy0 + 1 : ph - 1;
    const int xx0 = (int)(px >> 16), fx = (int)((px & 0xffff) >> 8), xx1 = xx0 + 1 < pw ? xx0 + 1 : pw - 1;
    d[(size_t)row * qw + x] = (uint8_t)vcp_bilerp(__ldg(s + (size_t)yy0 * pw + xx0), __ldg(s + (size_t)yy0 * pw + xx1),
                                                  __ldg(s + (size_t)yy1 * pw + xx0), __ldg(s + (size_t)yy1 * pw + xx1), fx, fy);
}

}  // namespace

void vcp_launch_k1_yuv420p(const uint8_t* in, size_t frame_bytes, int n0, int n, const VcpGeom& g,
                           const VcpBufs& b, cudaStream_t st) {
    if (n <= 0) return;
    int tiles = g.ys / 32;
    dim3 grid((tiles + 63) / 64, (g.ch + 2 * VCP_PAD) / 2, n);
    k1_yuv420p_kernel<<<grid, 64, 0, st>>>(in, frame_bytes, n0, g, b);
}

void vcp_launch_k1_to_yuv420p(const uint8_t* in, size_t in_fb, int fmt, int w, int h, uint8_t* out, size_t out_fb, int n,
                              cudaStream_t st) {
    if (n <= 0) return;
    const int cw = (w + 1) / 2, chh = (h + 1) / 2;
    dim3 grid((cw + 127) / 128, chh, n);
    k1_to_yuv420p_kernel<<<grid, 128, 0, st>>>(in, in_fb, fmt, w, h, out, out_fb);
}

void vcp_launch_k1_scale(const uint8_t* in, size_t in_fb, int sw, int sh, uint8_t* out, size_t out_fb, int dw, int dh,
                         int n, cudaStream_t st) {
    if (n <= 0) return;
    dim3 grid((dw + 127) / 128, dh + 2 * ((dh + 1) / 2), n);
    k1_scale_kernel<<<grid, 128, 0, st>>>(in, in_fb, sw, sh, out, out_fb, dw, dh);
}
