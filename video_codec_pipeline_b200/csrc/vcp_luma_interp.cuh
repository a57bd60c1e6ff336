// Luma sub-pel prediction (H.264 8.4.2.2.1) from the per-picture half-sample planes.
//
// Every quarter-sample position is either one sample of the half-sample grid {G, b, h, j} or the
// rounded average (__vavgu4, exactly (a+b+1)>>1) of two of them:
//   both coordinates even (in quarter units)  : the grid sample itself
//   one coordinate odd                        : the two grid neighbours along that axis
//   both odd                                  : the diagonal pair (x odd, y even) + (x even, y odd)
//                                               in half-sample units (e = b+h, g = b+m, p = h+s, r = m+s)
// The planes are built once per reconstructed picture by k2_hpel.cu, so a prediction costs two
// unaligned 8-byte fetches per lane and no filtering.
#ifndef VCP_LUMA_INTERP_CUH
#define VCP_LUMA_INTERP_CUH

#include "vcp_dev.cuh"

// the two grid samples of a quarter-sample displacement, in half-sample units
struct HpelPoints { int x1, y1, x2, y2; };

__device__ __forceinline__ HpelPoints hpel_points(int qx, int qy) {
    HpelPoints h;
    const int xa = (qx - 1) >> 1, ya = (qy - 1) >> 1;
    if (!(qx & 1) && !(qy & 1)) { h.x1 = h.x2 = qx >> 1; h.y1 = h.y2 = qy >> 1; }
    else if (!(qy & 1)) { h.x1 = xa; h.x2 = xa + 1; h.y1 = h.y2 = qy >> 1; }
    else if (!(qx & 1)) { h.x1 = h.x2 = qx >> 1; h.y1 = ya; h.y2 = ya + 1; }
    else {
        h.x1 = (xa & 1) ? xa : xa + 1; h.x2 = (xa & 1) ? xa + 1 : xa;    // x1 odd, x2 even
        h.y1 = (ya & 1) ? ya + 1 : ya; h.y2 = (ya & 1) ? ya : ya + 1;    // y1 even, y2 odd
    }
    return h;
}
__device__ __forceinline__ int hpel_plane(int x, int y) { return (x & 1) | ((y & 1) << 1); }   // 0 G, 1 B, 2 H, 3 J

// ---- straight from the planes in global memory (motion compensation: one fetch per macroblock)
// `blk`: pointer into the G plane at this lane's row / column for displacement (0,0)
__device__ __forceinline__ uint2 hpel_fetch8(const uint8_t* __restrict__ blk, int qx, int qy, int ys, size_t ysize) {
    const HpelPoints h = hpel_points(qx, qy);
    const uint2 a = ld8_unaligned(blk + hpel_plane(h.x1, h.y1) * ysize + (ptrdiff_t)(h.y1 >> 1) * ys + (h.x1 >> 1));
    if (h.x1 == h.x2 && h.y1 == h.y2) return a;
    const uint2 c = ld8_unaligned(blk + hpel_plane(h.x2, h.y2) * ysize + (ptrdiff_t)(h.y2 >> 1) * ys + (h.x2 >> 1));
    return make_uint2(__vavgu4(a.x, c.x), __vavgu4(a.y, c.y));
}

// ---- from a window staged in shared memory (sub-pel search: 16 candidates per macroblock) ----
// 18 rows (y = -1..16) of 6 aligned words per plane, covering x = -1..16 around an integer centre
struct __align__(16) HpelWindow {
    uint32_t w[4][18][6];
};

// centre: G-plane pointer of the block's sample (0,0) at the integer centre.  Returns the byte
// misalignment of column -1 inside the first staged word.
__device__ __forceinline__ int hpel_window_stage(HpelWindow& W, const uint8_t* __restrict__ centre, int ys, size_t ysize,
                                                 int lane, int nplanes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(centre - 1);
    const int mis = (int)(a & 3);
    const uint8_t* base = centre - 1 - mis - ys;       // aligned word holding (x=-1, y=-1)
    // 18 rows x 6 words per plane: lane -> word (lane & 7) < 6, rows (lane >> 3) + 4k
    const int c = lane & 7, r0 = lane >> 3;
    if (c < 6) {
        for (int p = 0; p < nplanes; p++) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(base + p * ysize) + c;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int r = r0 + 4 * k;
                if (r < 18) W.w[p][r][c] = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(src) + (ptrdiff_t)r * ys));
            }
        }
    }
    __syncwarp();
    return mis;
}

// 8 bytes at byte offset `off` of a word row in shared memory
__device__ __forceinline__ uint2 hpel_row8(const uint32_t* __restrict__ r, int off) {
    const int wi = off >> 2;
    const uint32_t sh = (uint32_t)(off & 3) * 8;
    const uint32_t w0 = r[wi], w1 = r[wi + 1], w2 = r[wi + 2];
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

// 8 samples of row `row`, columns hx.., at quarter displacement (qx,qy) in [-4,4) from the centre;
// `lane_off` = mis + 1 + hx (byte offset of this lane's first sample at displacement 0)
__device__ __forceinline__ uint2 hpel_window_fetch8(const HpelWindow& W, int lane_off, int qx, int qy, int row) {
    const HpelPoints h = hpel_points(qx, qy);
    const uint2 a = hpel_row8(W.w[hpel_plane(h.x1, h.y1)][row + (h.y1 >> 1) + 1], lane_off + (h.x1 >> 1));
    if (h.x1 == h.x2 && h.y1 == h.y2) return a;
    const uint2 c = hpel_row8(W.w[hpel_plane(h.x2, h.y2)][row + (h.y2 >> 1) + 1], lane_off + (h.x2 >> 1));
    return make_uint2(__vavgu4(a.x, c.x), __vavgu4(a.y, c.y));
}

#endif  // VCP_LUMA_INTERP_CUH
