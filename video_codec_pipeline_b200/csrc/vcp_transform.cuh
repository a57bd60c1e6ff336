// 4x4 integer transform / quantisation helpers (H.264 8.5), register-resident per thread.
#ifndef VCP_TRANSFORM_CUH
#define VCP_TRANSFORM_CUH

#include "vcp_dev.cuh"

#define VCP_TAB static __device__ const
#include "h264_tables.h"

__device__ __forceinline__ void vcp_fdct4(const int d[16], int w[16]) {
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int a0 = d[4 * i] + d[4 * i + 3], a1 = d[4 * i + 1] + d[4 * i + 2];
        const int a2 = d[4 * i + 1] - d[4 * i + 2], a3 = d[4 * i] - d[4 * i + 3];
        t[4 * i] = a0 + a1; t[4 * i + 1] = 2 * a3 + a2; t[4 * i + 2] = a0 - a1; t[4 * i + 3] = a3 - 2 * a2;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int a0 = t[i] + t[12 + i], a1 = t[4 + i] + t[8 + i];
        const int a2 = t[4 + i] - t[8 + i], a3 = t[i] - t[12 + i];
        w[i] = a0 + a1; w[4 + i] = 2 * a3 + a2; w[8 + i] = a0 - a1; w[12 + i] = a3 - 2 * a2;
    }
}

__device__ __forceinline__ void vcp_idct4(const int c[16], int r[16]) {
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int e0 = c[4 * i] + c[4 * i + 2], e1 = c[4 * i] - c[4 * i + 2];
        const int e2 = (c[4 * i + 1] >> 1) - c[4 * i + 3], e3 = c[4 * i + 1] + (c[4 * i + 3] >> 1);
        t[4 * i] = e0 + e3; t[4 * i + 1] = e1 + e2; t[4 * i + 2] = e1 - e2; t[4 * i + 3] = e0 - e3;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int e0 = t[i] + t[8 + i], e1 = t[i] - t[8 + i];
        const int e2 = (t[4 + i] >> 1) - t[12 + i], e3 = t[4 + i] + (t[12 + i] >> 1);
        r[i] = (e0 + e3 + 32) >> 6; r[4 + i] = (e1 + e2 + 32) >> 6;
        r[8 + i] = (e1 - e2 + 32) >> 6; r[12 + i] = (e0 - e3 + 32) >> 6;
    }
}

__device__ __forceinline__ int vcp_quant1(int w, int mf, int f, int qbits) {
    const int a = w < 0 ? -w : w;
    int l = (int)(((long long)a * mf + f) >> qbits);
    l = l > 2047 ? 2047 : l;
    return w < 0 ? -l : l;
}

// Quantise raster coefficients w into zig-zag levels lv (positions first..15), dequantise
// into raster c (positions < first are left untouched).  Returns the non-zero count.
__device__ __forceinline__ int vcp_quant_dequant4x4(const int w[16], int qp, bool intra, int first, int lv[16], int c[16]) {
    const int qbits = 15 + qp / 6, f = (1 << qbits) / (intra ? 3 : 6), rem = qp % 6, sh = qp / 6;
    const int mf0 = vcp_quant_mf[rem][0], mf1 = vcp_quant_mf[rem][1], mf2 = vcp_quant_mf[rem][2];
    const int v0 = vcp_dequant_v[rem][0], v1 = vcp_dequant_v[rem][1], v2 = vcp_dequant_v[rem][2];
    int nz = 0;
    constexpr int zz[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
    constexpr int cls[16] = {0, 2, 0, 2, 2, 1, 2, 1, 0, 2, 0, 2, 2, 1, 2, 1};
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k < first) { lv[k] = 0; continue; }
        const int i = zz[k];
        const int mf = cls[i] == 0 ? mf0 : (cls[i] == 1 ? mf1 : mf2);
        const int v = cls[i] == 0 ? v0 : (cls[i] == 1 ? v1 : v2);
        const int l = vcp_quant1(w[i], mf, f, qbits);
        lv[k] = l;
        nz += l != 0;
        c[i] = (l * v) << sh;
    }
    return nz;
}

// store 16 levels (int) as int16 to a 32-byte aligned record
__device__ __forceinline__ void vcp_store_levels16(int16_t* dst, const int lv[16]) {
    uint4 a, b;
    a.x = (uint32_t)(lv[0] & 0xffff) | ((uint32_t)lv[1] << 16);
    a.y = (uint32_t)(lv[2] & 0xffff) | ((uint32_t)lv[3] << 16);
    a.z = (uint32_t)(lv[4] & 0xffff) | ((uint32_t)lv[5] << 16);
    a.w = (uint32_t)(lv[6] & 0xffff) | ((uint32_t)lv[7] << 16);
    b.x = (uint32_t)(lv[8] & 0xffff) | ((uint32_t)lv[9] << 16);
    b.y = (uint32_t)(lv[10] & 0xffff) | ((uint32_t)lv[11] << 16);
    b.z = (uint32_t)(lv[12] & 0xffff) | ((uint32_t)lv[13] << 16);
    b.w = (uint32_t)(lv[14] & 0xffff) | ((uint32_t)lv[15] << 16);
    reinterpret_cast<uint4*>(dst)[0] = a;
    reinterpret_cast<uint4*>(dst)[1] = b;
}

// pack 4 clipped samples pred+res into one word
__device__ __forceinline__ uint32_t vcp_recon4(uint32_t pred, int r0, int r1, int r2, int r3) {
    const uint32_t a = (uint32_t)vcp_clip255((int)(pred & 255) + r0);
    const uint32_t b = (uint32_t)vcp_clip255((int)((pred >> 8) & 255) + r1);
    const uint32_t c = (uint32_t)vcp_clip255((int)((pred >> 16) & 255) + r2);
    const uint32_t d = (uint32_t)vcp_clip255((int)(pred >> 24) + r3);
    return a | (b << 8) | (c << 16) | (d << 24);
}

// Chroma DC of one plane held by 4 consecutive lanes (lane&3 = block index, raster 2x2):
// forward 2x2 Hadamard, quantise, inverse Hadamard, scale.  Returns the level of this
// lane's position; `deq` receives the dequantised DC of this lane's block.
__device__ __forceinline__ int vcp_chroma_dc(int dc, int qpc, bool intra, int lane, int& deq) {
    const int base = lane & ~3, i = lane & 3;
    const int d0 = __shfl_sync(0xffffffffu, dc, base), d1 = __shfl_sync(0xffffffffu, dc, base + 1);
    const int d2 = __shfl_sync(0xffffffffu, dc, base + 2), d3 = __shfl_sync(0xffffffffu, dc, base + 3);
    const int s1 = (i & 1) ? -1 : 1, s2 = (i & 2) ? -1 : 1;
    const int h = d0 + s1 * d1 + s2 * d2 + s1 * s2 * d3;
    const int qbits = 15 + qpc / 6, f = (1 << qbits) / (intra ? 3 : 6);
    const int l = vcp_quant1(h, vcp_quant_mf[qpc % 6][0], 2 * f, qbits + 1);
    const int c0 = __shfl_sync(0xffffffffu, l, base), c1 = __shfl_sync(0xffffffffu, l, base + 1);
    const int c2 = __shfl_sync(0xffffffffu, l, base + 2), c3 = __shfl_sync(0xffffffffu, l, base + 3);
    const int gq = c0 + s1 * c1 + s2 * c2 + s1 * s2 * c3;
    deq = ((gq * 16 * vcp_dequant_v[qpc % 6][0]) << (qpc / 6)) >> 5;
    return l;
}

#endif  // VCP_TRANSFORM_CUH
