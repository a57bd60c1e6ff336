// K2c — half-sample planes of a reconstructed picture (H.264 8.4.2.2.1: samples b, h, j).
//
// The luma six-tap filter is the expensive part of sub-pel motion search and of motion
// compensation.  Computing it per macroblock (with a 5-sample halo on a 16x16 block) costs
// ~2.5x the samples and sat on the frame-to-frame chain twice (refine and reconstruction).
// Instead every reconstructed picture gets its three half-sample planes computed ONCE, right
// after deblocking and border extension, in one streaming pass; every quarter-sample
// prediction is then one plane sample or the rounded average of two (vcp_hpel.cuh).
//
// Layout: a reconstruction slot holds four consecutive luma planes of identical geometry
//   G (integer samples), B (x+1/2), H (y+1/2), J (x+1/2, y+1/2)        -> vcp_rec_luma()
// B[y][x] sits between G[y][x] and G[y][x+1]; H[y][x] between G[y][x] and G[y+1][x].
//
// Tile = 64 x 32 samples per CTA of 128 threads; the integer tile (+halo) and the unclipped
// horizontal sums b1 live in shared memory; horizontal taps use dp4a (u8 x s8), vertical taps
// run on four rows per thread so the six-row window is loaded 9/4 times instead of 6.
// HBM traffic: read 1 plane, write 3 planes (per luma sample: 1 + 3 bytes).
#include "vcp_dev.cuh"

namespace {

constexpr int HT_W = 64, HT_H = 32, HT_ROWS = HT_H + 5, HT_GW = 72;   // G tile: rows Y0-2..Y0+34, cols X0-4..X0+67

__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int tap6i(int a, int b, int c, int d, int e, int f) { return (a + f) + 20 * (c + d) - 5 * (b + e); }
__device__ __forceinline__ uint32_t clip_pack4(int v0, int v1, int v2, int v3, int rnd, int sh) {
    const uint32_t a = (uint32_t)vcp_clip255((v0 + rnd) >> sh), b = (uint32_t)vcp_clip255((v1 + rnd) >> sh);
    const uint32_t c = (uint32_t)vcp_clip255((v2 + rnd) >> sh), d = (uint32_t)vcp_clip255((v3 + rnd) >> sh);
    return a | (b << 8) | (c << 16) | (d << 24);
}

// grid: x = column tiles, y = row tiles, z = GOP of the group
__global__ void __launch_bounds__(128) hpel_planes_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ __align__(16) uint32_t Gs[HT_ROWS][HT_GW / 4];
    __shared__ __align__(16) uint32_t b1s[HT_ROWS][HT_W / 2];   // int16 pairs
    const int gi = blockIdx.z + s.g0;
    const int slot = vcp_rec_slot(s, gi, s.t);
    uint8_t* G = vcp_rec_luma(b, g, slot);
    uint8_t* PB = G + g.ysize;
    uint8_t* PH = G + 2 * g.ysize;
    uint8_t* PJ = G + 3 * g.ysize;
    const int rows = g.ch + 2 * VCP_PAD, wwords = g.ys / 4;
    const int X0 = blockIdx.x * HT_W, Y0 = blockIdx.y * HT_H;
    const int tid = threadIdx.x;

    for (int i = tid; i < HT_ROWS * (HT_GW / 4); i += 128) {
        const int r = i / (HT_GW / 4), c = i % (HT_GW / 4);
        const int y = vcp_clip3(0, rows - 1, Y0 - 2 + r);
        const int wc = vcp_clip3(0, wwords - 1, X0 / 4 - 1 + c);
        Gs[r][c] = __ldg(reinterpret_cast<const uint32_t*>(G + (size_t)y * g.ys) + wc);
    }
    __syncthreads();
    // horizontal pass: b1 for all 37 rows, B for the 32 output rows
    constexpr int C0 = 0x1414FB01;            // (1, -5, 20, 20) little-endian s8
    constexpr int C1 = 0x000001FB;            // (-5, 1, 0, 0)
    for (int i = tid; i < HT_ROWS * (HT_W / 4); i += 128) {
        const int r = i >> 4, cg = i & 15;
        const uint32_t w0 = Gs[r][cg], w1 = Gs[r][cg + 1], w2 = Gs[r][cg + 2];
        // output k uses bytes k+2 .. k+7 of (w0,w1,w2)
        const uint32_t a0 = __funnelshift_r(w0, w1, 16), a1 = __funnelshift_r(w0, w1, 24), a2 = w1, a3 = __funnelshift_r(w1, w2, 8);
        const uint32_t e0 = __funnelshift_r(w1, w2, 16), e1 = __funnelshift_r(w1, w2, 24), e2 = w2, e3 = w2 >> 8;
        const int v0 = dp4a_us(e0, C1, dp4a_us(a0, C0, 0));
        const int v1 = dp4a_us(e1, C1, dp4a_us(a1, C0, 0));
        const int v2 = dp4a_us(e2, C1, dp4a_us(a2, C0, 0));
        const int v3 = dp4a_us(e3, C1, dp4a_us(a3, C0, 0));
        b1s[r][2 * cg] = (uint32_t)(v0 & 0xffff) | ((uint32_t)v1 << 16);
        b1s[r][2 * cg + 1] = (uint32_t)(v2 & 0xffff) | ((uint32_t)v3 << 16);
        const int y = Y0 + r - 2;
        if (r >= 2 && r < 2 + HT_H && y < rows)
            *reinterpret_cast<uint32_t*>(PB + (size_t)y * g.ys + X0 + 4 * cg) = clip_pack4(v0, v1, v2, v3, 16, 5);
    }
    __syncthreads();
    // vertical pass: thread = 4 columns x 4 rows; output row y uses tile rows y .. y+5
    {
        const int cg = tid & 15, ry = (tid >> 4) * 4;
        int col[9][4];
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const uint32_t w = Gs[ry + k][cg + 1];
            col[k][0] = w & 255; col[k][1] = (w >> 8) & 255; col[k][2] = (w >> 16) & 255; col[k][3] = w >> 24;
        }
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int y = Y0 + ry + o;
            if (y < rows) {
                int v[4];
#pragma unroll
                for (int x = 0; x < 4; x++) v[x] = tap6i(col[o][x], col[o + 1][x], col[o + 2][x], col[o + 3][x], col[o + 4][x], col[o + 5][x]);
                *reinterpret_cast<uint32_t*>(PH + (size_t)y * g.ys + X0 + 4 * cg) = clip_pack4(v[0], v[1], v[2], v[3], 16, 5);
            }
        }
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const uint32_t w0 = b1s[ry + k][2 * cg], w1 = b1s[ry + k][2 * cg + 1];
            col[k][0] = (int)(int16_t)(w0 & 0xffff); col[k][1] = (int)w0 >> 16;
            col[k][2] = (int)(int16_t)(w1 & 0xffff); col[k][3] = (int)w1 >> 16;
        }
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int y = Y0 + ry + o;
            if (y < rows) {
                int v[4];
#pragma unroll
                for (int x = 0; x < 4; x++) v[x] = tap6i(col[o][x], col[o + 1][x], col[o + 2][x], col[o + 3][x], col[o + 4][x], col[o + 5][x]);
                *reinterpret_cast<uint32_t*>(PJ + (size_t)y * g.ys + X0 + 4 * cg) = clip_pack4(v[0], v[1], v[2], v[3], 512, 10);
            }
        }
    }
}

// ---- HEVC (8.5.3.3.3.1): the same three planes with the 8-tap half-sample filter (-1, 4, -11, 40, 40, -11, 4, -1).
// B and H: (sum + 32) >> 6; J: ((vertical sum of the UNSHIFTED horizontal sums) >> 6, then + 32) >> 6 -- the
// standard's two-stage rounding for 8-bit video (shift1 = 0, shift2 = 6, then the default weighted prediction).
// Same tile as above; the filter reaches one sample further on each side: G rows Y0-3..Y0+35, cols X0-4..X0+67.
constexpr int VT_ROWS = HT_H + 7;
__device__ __forceinline__ int tap8i(int a, int b, int c, int d, int e, int f, int g, int h) {
    return 40 * (d + e) - 11 * (c + f) + 4 * (b + g) - (a + h);
}

__global__ void __launch_bounds__(128) hevc_hpel_planes_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ __align__(16) uint32_t Gs[VT_ROWS][HT_GW / 4];
    __shared__ __align__(16) uint32_t b1s[VT_ROWS][HT_W / 2];   // int16 pairs: unshifted horizontal sums
    const int gi = blockIdx.z + s.g0;
    const int slot = vcp_rec_slot(s, gi, s.t);
    uint8_t* G = vcp_rec_luma(b, g, slot);
    uint8_t* PB = G + g.ysize;
    uint8_t* PH = G + 2 * g.ysize;
    uint8_t* PJ = G + 3 * g.ysize;
    const int rows = g.ch + 2 * VCP_PAD, wwords = g.ys / 4;
    const int X0 = blockIdx.x * HT_W, Y0 = blockIdx.y * HT_H;
    const int tid = threadIdx.x;

    for (int i = tid; i < VT_ROWS * (HT_GW / 4); i += 128) {
        const int r = i / (HT_GW / 4), c = i % (HT_GW / 4);
        const int y = vcp_clip3(0, rows - 1, Y0 - 3 + r);
        const int wc = vcp_clip3(0, wwords - 1, X0 / 4 - 1 + c);
        Gs[r][c] = __ldg(reinterpret_cast<const uint32_t*>(G + (size_t)y * g.ys) + wc);
    }
    __syncthreads();
    // horizontal pass: unshifted sums for all 39 rows, B for the 32 output rows
    constexpr int C0 = 0x28F504FF;            // (-1, 4, -11, 40) little-endian s8
    constexpr int C1 = (int)0xFF04F528;       // (40, -11, 4, -1)
    for (int i = tid; i < VT_ROWS * (HT_W / 4); i += 128) {
        const int r = i >> 4, cg = i & 15;
        const uint32_t w0 = Gs[r][cg], w1 = Gs[r][cg + 1], w2 = Gs[r][cg + 2];
        // output k (sample X0 + 4 cg + k + 1/2) uses tile bytes 4 cg + k + 1 .. 4 cg + k + 8
        const uint32_t a0 = __funnelshift_r(w0, w1, 8), a1 = __funnelshift_r(w0, w1, 16), a2 = __funnelshift_r(w0, w1, 24), a3 = w1;
        const uint32_t e0 = __funnelshift_r(w1, w2, 8), e1 = __funnelshift_r(w1, w2, 16), e2 = __funnelshift_r(w1, w2, 24), e3 = w2;
        const int v0 = dp4a_us(e0, C1, dp4a_us(a0, C0, 0));
        const int v1 = dp4a_us(e1, C1, dp4a_us(a1, C0, 0));
        const int v2 = dp4a_us(e2, C1, dp4a_us(a2, C0, 0));
        const int v3 = dp4a_us(e3, C1, dp4a_us(a3, C0, 0));
        b1s[r][2 * cg] = (uint32_t)(v0 & 0xffff) | ((uint32_t)v1 << 16);
        b1s[r][2 * cg + 1] = (uint32_t)(v2 & 0xffff) | ((uint32_t)v3 << 16);
        const int y = Y0 + r - 3;
        if (r >= 3 && r < 3 + HT_H && y < rows)
            *reinterpret_cast<uint32_t*>(PB + (size_t)y * g.ys + X0 + 4 * cg) = clip_pack4(v0, v1, v2, v3, 32, 6);
    }
    __syncthreads();
    // vertical pass: thread = 4 columns x 4 rows; output row y uses tile rows y .. y+7
    {
        const int cg = tid & 15, ry = (tid >> 4) * 4;
        int col[11][4];
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const uint32_t w = Gs[ry + k][cg + 1];
            col[k][0] = w & 255; col[k][1] = (w >> 8) & 255; col[k][2] = (w >> 16) & 255; col[k][3] = w >> 24;
        }
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int y = Y0 + ry + o;
            if (y < rows) {
                int v[4];
#pragma unroll
                for (int x = 0; x < 4; x++)
                    v[x] = tap8i(col[o][x], col[o + 1][x], col[o + 2][x], col[o + 3][x], col[o + 4][x], col[o + 5][x], col[o + 6][x], col[o + 7][x]);
                *reinterpret_cast<uint32_t*>(PH + (size_t)y * g.ys + X0 + 4 * cg) = clip_pack4(v[0], v[1], v[2], v[3], 32, 6);
            }
        }
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const uint32_t w0 = b1s[ry + k][2 * cg], w1 = b1s[ry + k][2 * cg + 1];
            col[k][0] = (int)(int16_t)(w0 & 0xffff); col[k][1] = (int)w0 >> 16;
            col[k][2] = (int)(int16_t)(w1 & 0xffff); col[k][3] = (int)w1 >> 16;
        }
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int y = Y0 + ry + o;
            if (y < rows) {
                int v[4];
#pragma unroll
                for (int x = 0; x < 4; x++)
                    v[x] = tap8i(col[o][x], col[o + 1][x], col[o + 2][x], col[o + 3][x], col[o + 4][x], col[o + 5][x], col[o + 6][x], col[o + 7][x]) >> 6;
                *reinterpret_cast<uint32_t*>(PJ + (size_t)y * g.ys + X0 + 4 * cg) = clip_pack4(v[0], v[1], v[2], v[3], 32, 6);
            }
        }
    }
}

}  // namespace

void vcp_launch_hpel(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    const int rows = g.ch + 2 * VCP_PAD;
    dim3 grid(g.ys / HT_W, (rows + HT_H - 1) / HT_H, s.ngop);
    if (g.hevc) hevc_hpel_planes_kernel<<<grid, 128, 0, st>>>(g, b, s);
    else hpel_planes_kernel<<<grid, 128, 0, st>>>(g, b, s);
}
