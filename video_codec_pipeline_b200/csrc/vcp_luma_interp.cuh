// Luma sub-pel interpolation (H.264 8.4.2.2.1) for one 16x16 macroblock, warp-cooperative.
//
// A warp builds, in shared memory, the integer window and the three half-sample planes
//   b (horizontal), h (vertical), j (centre)
// around an integer-pel centre.  Every quarter-pel prediction within (-1,+1) pel of the
// centre is then either one plane sample or the rounded average (__vavgu4, exactly
// (a+b+1)>>1) of two of them, so the 16 sub-pel candidates of the refine and the final
// prediction cost a few shared loads each.
#ifndef VCP_LUMA_INTERP_CUH
#define VCP_LUMA_INTERP_CUH

#include "vcp_dev.cuh"

struct __align__(16) LumaPlanes {
    uint8_t G[22][24];   // integer samples, rows y=-3..18 (idx y+3), cols x=-4..19 (idx x+4)
    int16_t b1[22][18];  // unclipped horizontal 6-tap sums, rows y=-3..18, x=-1..15 (idx x+1)
    uint8_t B[18][20];   // b, rows y=-1..16 (idx y+1), x=-1..15 (idx x+1)
    uint8_t H[18][24];   // h, rows y=-1..15 (idx y+1), x=-1..16 (idx x+1)
    uint8_t J[18][20];   // j, rows y=-1..15 (idx y+1), x=-1..15 (idx x+1)
};

__device__ __forceinline__ int vcp_tap6(int a, int b, int c, int d, int e, int f) {
    return a - 5 * b + 20 * c + 20 * d - 5 * e + f;
}

// centre: pointer to the integer sample (0,0) of the block in the reference plane.
// Lane mappings are fixed (no div/mod in the loops): a lane owns one column (or word) and
// strides over rows.
__device__ __forceinline__ void luma_planes_build(LumaPlanes& P, const uint8_t* __restrict__ centre, int stride,
                                                  int lane, bool needB, bool needH, bool needJ) {
    {   // integer window: 22 rows x 6 words; lane -> word (lane&7) < 6, rows (lane>>3) + 4k
        const int c = lane & 7, r0 = lane >> 3;
        if (c < 6) {
            const uint8_t* src = centre - 4 + 4 * c + (ptrdiff_t)(r0 - 3) * stride;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const int r = r0 + 4 * k;
                if (r < 22) reinterpret_cast<uint32_t*>(&P.G[r][0])[c] = ld4_unaligned(src + (ptrdiff_t)(4 * k) * stride);
            }
        }
    }
    __syncwarp();
    if (needB || needJ) {
        // b1: 22 rows x 17 columns.  lanes 0-15 / 16-31 take even / odd rows of columns 0..15;
        // column 16 is done by lanes 0..21 (one row each)
        const int x = lane & 15, r0 = lane >> 4;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const int r = r0 + 2 * k;
            const uint8_t* p = &P.G[r][x + 1];
            P.b1[r][x] = (int16_t)vcp_tap6(p[0], p[1], p[2], p[3], p[4], p[5]);
        }
        if (lane < 22) {
            const uint8_t* p = &P.G[lane][17];
            P.b1[lane][16] = (int16_t)vcp_tap6(p[0], p[1], p[2], p[3], p[4], p[5]);
        }
    }
    if (needH) {
        // h: 17 rows x 18 columns; lanes 0-15 / 16-31 take rows of columns 0..15 alternately,
        // columns 16,17 by lanes 0..16 / 17..(unused) in a second pass
        const int x = lane & 15, r0 = lane >> 4;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const int yy = r0 + 2 * k;
            if (yy < 17) {
                const int v = vcp_tap6(P.G[yy][x + 3], P.G[yy + 1][x + 3], P.G[yy + 2][x + 3], P.G[yy + 3][x + 3],
                                       P.G[yy + 4][x + 3], P.G[yy + 5][x + 3]);
                P.H[yy][x] = (uint8_t)vcp_clip255((v + 16) >> 5);
            }
        }
#pragma unroll
        for (int xx = 16; xx < 18; xx++) {
            if (lane < 17) {
                const int v = vcp_tap6(P.G[lane][xx + 3], P.G[lane + 1][xx + 3], P.G[lane + 2][xx + 3], P.G[lane + 3][xx + 3],
                                       P.G[lane + 4][xx + 3], P.G[lane + 5][xx + 3]);
                P.H[lane][xx] = (uint8_t)vcp_clip255((v + 16) >> 5);
            }
        }
    }
    __syncwarp();
    if (needB) {
        const int x = lane & 15, r0 = lane >> 4;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const int yy = r0 + 2 * k;
            P.B[yy][x] = (uint8_t)vcp_clip255((P.b1[yy + 2][x] + 16) >> 5);
        }
        if (lane < 18) P.B[lane][16] = (uint8_t)vcp_clip255((P.b1[lane + 2][16] + 16) >> 5);
    }
    if (needJ) {
        const int x = lane & 15, r0 = lane >> 4;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const int yy = r0 + 2 * k;
            if (yy < 17) {
                const int v = vcp_tap6(P.b1[yy][x], P.b1[yy + 1][x], P.b1[yy + 2][x], P.b1[yy + 3][x], P.b1[yy + 4][x],
                                       P.b1[yy + 5][x]);
                P.J[yy][x] = (uint8_t)vcp_clip255((v + 512) >> 10);
            }
        }
        if (lane < 17) {
            const int v = vcp_tap6(P.b1[lane][16], P.b1[lane + 1][16], P.b1[lane + 2][16], P.b1[lane + 3][16],
                                   P.b1[lane + 4][16], P.b1[lane + 5][16]);
            P.J[lane][16] = (uint8_t)vcp_clip255((v + 512) >> 10);
        }
    }
    __syncwarp();
}

// which planes does the fractional position (fx,fy) need?
__device__ __forceinline__ void luma_planes_needs(int fx, int fy, bool& needB, bool& needH, bool& needJ) {
    const int c = fy * 4 + fx;
    // b: 1,2,3,5,6,7,13,14,15   h: 4,5,7,8,9,11,12,13,15   j: 6,9,10,11,14
    needB = (0xE0EEu >> c) & 1;
    needH = (0xBBB0u >> c) & 1;
    needJ = (0x4E40u >> c) & 1;
}

// 8 prediction samples of row `row`, columns hx..hx+7, at quarter-pel offset (qx,qy) from the
// centre; qx,qy in [-4,3].
__device__ __forceinline__ uint2 luma_planes_fetch8(const LumaPlanes& P, int qx, int qy, int row, int hx) {
    const int ix = qx >> 2, iy = qy >> 2, fx = qx & 3, fy = qy & 3;
    const int y = iy + row, x = ix + hx;
#define PG(dx, dy) ld8_unaligned(&P.G[y + (dy) + 3][x + (dx) + 4])
#define PB(dy) ld8_unaligned(&P.B[y + (dy) + 1][x + 1])
#define PH(dx) ld8_unaligned(&P.H[y + 1][x + (dx) + 1])
#define PJ() ld8_unaligned(&P.J[y + 1][x + 1])
    uint2 a, c;
    switch (fy * 4 + fx) {
    case 0: return PG(0, 0);
    case 1: a = PG(0, 0); c = PB(0); break;
    case 2: return PB(0);
    case 3: a = PG(1, 0); c = PB(0); break;
    case 4: a = PG(0, 0); c = PH(0); break;
    case 5: a = PB(0); c = PH(0); break;
    case 6: a = PB(0); c = PJ(); break;
    case 7: a = PB(0); c = PH(1); break;
    case 8: return PH(0);
    case 9: a = PH(0); c = PJ(); break;
    case 10: return PJ();
    case 11: a = PJ(); c = PH(1); break;
    case 12: a = PG(0, 1); c = PH(0); break;
    case 13: a = PH(0); c = PB(1); break;
    case 14: a = PJ(); c = PB(1); break;
    default: a = PH(1); c = PB(1); break;
    }
#undef PG
#undef PB
#undef PH
#undef PJ
    return make_uint2(__vavgu4(a.x, c.x), __vavgu4(a.y, c.y));
}

#endif  // VCP_LUMA_INTERP_CUH
