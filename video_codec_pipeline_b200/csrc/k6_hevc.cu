// K6 — HEVC (ITU-T H.265) reconstruction kernels: the h265-* presets of the reference
// (/root/reference/internal/config/config.go:47-50) on the same execution model as the H.264 path:
// closed GOPs in lock-step, motion search shared with H.264 (pre-pass on originals + refine on the
// reconstruction), everything else coding-unit-parallel, intra CUs on an anti-diagonal wavefront.
//
// Stream structure (oracle/hevc_oracle.inc.c restates the same): Main profile, CTB = CU = 16x16 in raster
// order, luma transform blocks 8x8 (the 16x16 root splits because MaxTb = 8), chroma 4x4, DCT only, DC intra
// prediction per transform block (every CU of an IDR picture; scene-cut CUs of P pictures), one 16x16 PU with
// full-sample or (hevc_subpel) half-sample luma vectors read from the 8-tap planes of k2_hpel.cu, chroma by the
// 4-tap filters, AMVP / merge (one candidate) / skip decided after all vectors exist (hevc_cuinfo_kernel), in-loop
// deblocking as two order-free passes (hevc_deblock_kernel), optional SAO (hevc_sao_kernel).
// Entropy coding: k5_cabac.cu (hevc_bins_kernel + the shared arithmetic coder).
//
//   hevc_p_recon_kernel : warp = CU; eight lanes per 8x8 luma transform block (a row, a column, a row each, through
//                         shared memory), then four lanes per 4x4 chroma block the same way
//   hevc_i_recon_kernel : CTA = (GOP, slice), wavefront over anti-diagonals, warp = CU, the four transform
//                         units of a CU in z-order (each predicts from the reconstruction of the previous ones);
//                         <FIX>: only the CUs of a P picture that the refine flagged intra
//   hevc_cuinfo_kernel  : thread = CU: merge candidate, AMVP list, vector difference, skip
//   hevc_deblock_kernel : thread = 4-line segment of an 8x8-grid edge; vertical edges, then horizontal edges
//   hevc_sao_kernel     : warp = CTB: edge-offset statistics on the deblocked picture, decision, filtered samples
//                         into the slot's scratch plane; hevc_sao_copy_kernel moves them back
//
// Record layout per CU (same arrays as H.264): mbtype 0 intra / 1 inter / 2 skip; cbp bits 0-3 = cbf_luma of
// the four TUs, bit 4 = merge_flag, bit 5 = mvp_l0_flag; modes bits 0-3 = cbf_cb, 4-7 = cbf_cr; levels:
// luma [z*64 + y*8 + x], Cb [256 + z*16 + y*4 + x], Cr [320 + z*16 + y*4 + x].
#include "vcp_dev.cuh"
#include "vcp_luma_interp.cuh"
#include "vcp_hevc_qpel.cuh"

#define VCP_TAB static __device__ const
#include "hevc_tables.h"

namespace {

constexpr int HV_T_STRIDE = 72;   // ints per 8x8 staging block (64 + padding: the four blocks land on different banks)

struct __align__(16) HvScratch {
    uint8_t pred[16][16];         // luma prediction, then reconstruction
    uint8_t cpred[2][8][8];
    int t[4][HV_T_STRIDE];
    int16_t lv[384];
};
static_assert(sizeof(int) * 4 * HV_T_STRIDE + sizeof(int16_t) * 384 >= sizeof(uint32_t) * HQ_WROWS * HQ_WPITCH, "the row-pass output aliases t + lv");

__device__ __forceinline__ int hevc_chroma_qp(int qp) { return qp < 30 ? qp : qp > 43 ? qp - 6 : hevc_qpc_tab[qp - 30]; }

// one-dimensional core transforms (8.6.4.2): exact integer sums, so the butterfly order is free
__device__ __forceinline__ void hv_fwd8(const int x[8], int y[8]) {
    const int e0 = x[0] + x[7], e1 = x[1] + x[6], e2 = x[2] + x[5], e3 = x[3] + x[4];
    const int o0 = x[0] - x[7], o1 = x[1] - x[6], o2 = x[2] - x[5], o3 = x[3] - x[4];
    const int ee0 = e0 + e3, eo0 = e0 - e3, ee1 = e1 + e2, eo1 = e1 - e2;
    y[0] = 64 * (ee0 + ee1); y[4] = 64 * (ee0 - ee1); y[2] = 83 * eo0 + 36 * eo1; y[6] = 36 * eo0 - 83 * eo1;
    y[1] = 89 * o0 + 75 * o1 + 50 * o2 + 18 * o3; y[3] = 75 * o0 - 18 * o1 - 89 * o2 - 50 * o3;
    y[5] = 50 * o0 - 89 * o1 + 18 * o2 + 75 * o3; y[7] = 18 * o0 - 50 * o1 + 75 * o2 - 89 * o3;
}
__device__ __forceinline__ void hv_inv8(const int c[8], int x[8]) {
    const int o0 = 89 * c[1] + 75 * c[3] + 50 * c[5] + 18 * c[7], o1 = 75 * c[1] - 18 * c[3] - 89 * c[5] - 50 * c[7];
    const int o2 = 50 * c[1] - 89 * c[3] + 18 * c[5] + 75 * c[7], o3 = 18 * c[1] - 50 * c[3] + 75 * c[5] - 89 * c[7];
    const int eo0 = 83 * c[2] + 36 * c[6], eo1 = 36 * c[2] - 83 * c[6], ee0 = 64 * (c[0] + c[4]), ee1 = 64 * (c[0] - c[4]);
    const int e0 = ee0 + eo0, e3 = ee0 - eo0, e1 = ee1 + eo1, e2 = ee1 - eo1;
    x[0] = e0 + o0; x[7] = e0 - o0; x[1] = e1 + o1; x[6] = e1 - o1; x[2] = e2 + o2; x[5] = e2 - o2; x[3] = e3 + o3; x[4] = e3 - o3;
}
__device__ __forceinline__ void hv_fwd4(const int x[4], int y[4]) {
    const int a = x[0] + x[3], b = x[1] + x[2], c = x[0] - x[3], d = x[1] - x[2];
    y[0] = 64 * (a + b); y[2] = 64 * (a - b); y[1] = 83 * c + 36 * d; y[3] = 36 * c - 83 * d;
}
__device__ __forceinline__ void hv_inv4(const int c[4], int x[4]) {
    const int e0 = 64 * (c[0] + c[2]), e1 = 64 * (c[0] - c[2]), o0 = 83 * c[1] + 36 * c[3], o1 = 36 * c[1] - 83 * c[3];
    x[0] = e0 + o0; x[3] = e0 - o0; x[1] = e1 + o1; x[2] = e1 - o1;
}
__device__ __forceinline__ int hv_quant1(int w, int qs, long long offs, int qbits) {
    const long long a = w < 0 ? -(long long)w : w;
    long long l = (a * qs + offs) >> qbits;
    if (l > 32767) l = 32767;
    return (int)(w < 0 ? -l : l);
}
__device__ __forceinline__ int hv_dequant1(int l, int ls, int sh, int bdshift) {
    const long long v = ((long long)l * 16 * ls) << sh;
    return vcp_clip3(-32768, 32767, (int)((v + (1 << (bdshift - 1))) >> bdshift));
}

// Luma 8x8 transform blocks, eight lanes per block (k = block slot, q = lane in the block: one row, then one
// column, then one row).  S.pred holds the prediction of the block at (bx,by) on entry and its reconstruction on
// exit; levels go to S.lv[k*64..].  Every lane of the warp calls this (the shared-memory passes are separated by
// warp barriers); `act` masks.  Returns the number of non-zero levels of the block (valid in its eight lanes).
__device__ __forceinline__ int hv_luma_tu(HvScratch& S, bool act, int k, int q, int bx, int by, const uint8_t* __restrict__ src,
                                          int stride, int qp, bool intra) {
    int* T = S.t[k];
    if (act) {   // row q: residual, horizontal transform, stage shift log2(8) + 8 - 9 = 2
        const uint2 s8 = *reinterpret_cast<const uint2*>(src + (size_t)q * stride);
        const uint2 p8 = *reinterpret_cast<const uint2*>(&S.pred[by + q][bx]);
        int d[8], y[8];
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const uint32_t sw = x < 4 ? s8.x : s8.y, pw = x < 4 ? p8.x : p8.y;
            d[x] = (int)((sw >> (8 * (x & 3))) & 255) - (int)((pw >> (8 * (x & 3))) & 255);
        }
        hv_fwd8(d, y);
#pragma unroll
        for (int x = 0; x < 8; x++) T[8 * q + ((x + q) & 7)] = (y[x] + 2) >> 2;     // rotated within the row: conflict-free column reads
    }
    __syncwarp();
    int nz = 0;
    if (act) {   // column q: vertical transform (shift 9), quantise, dequantise, vertical inverse (columns first, 8.6.4.2)
        const int qbits = 14 + qp / 6 + 4, qs = hevc_quant_scale[qp % 6], ls = hevc_level_scale[qp % 6], sh = qp / 6;
        const long long offs = (long long)(intra ? 171 : 85) << (qbits - 9);
        int in[8], w[8], c[8], g[8];
#pragma unroll
        for (int r = 0; r < 8; r++) in[r] = T[8 * r + ((q + r) & 7)];
        hv_fwd8(in, w);
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int l = hv_quant1((w[r] + 256) >> 9, qs, offs, qbits);
            S.lv[k * 64 + 8 * r + q] = (int16_t)l;
            nz += l != 0;
            c[r] = hv_dequant1(l, ls, sh, 6);
        }
        hv_inv8(c, g);
#pragma unroll
        for (int r = 0; r < 8; r++) T[8 * r + ((q + r) & 7)] = vcp_clip3(-32768, 32767, (g[r] + 64) >> 7);
    }
    nz += __shfl_xor_sync(0xffffffffu, nz, 1);
    nz += __shfl_xor_sync(0xffffffffu, nz, 2);
    nz += __shfl_xor_sync(0xffffffffu, nz, 4);
    __syncwarp();
    if (act && nz) {   // row q: horizontal inverse (shift 12), reconstruct in place of the prediction
        int in[8], o[8];
#pragma unroll
        for (int x = 0; x < 8; x++) in[x] = T[8 * q + ((x + q) & 7)];
        hv_inv8(in, o);
        const uint2 p8 = *reinterpret_cast<const uint2*>(&S.pred[by + q][bx]);
        uint2 r8 = make_uint2(0, 0);
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const uint32_t pw = x < 4 ? p8.x : p8.y;
            const uint32_t v = (uint32_t)vcp_clip255((int)((pw >> (8 * (x & 3))) & 255) + ((o[x] + 2048) >> 12)) << (8 * (x & 3));
            if (x < 4) r8.x |= v; else r8.y |= v;
        }
        *reinterpret_cast<uint2*>(&S.pred[by + q][bx]) = r8;
    }
    __syncwarp();
    return nz;
}

// Eight 4x4 chroma transform blocks of a coding unit at once, four lanes per block (blk = lane >> 2 = plane * 4 + z,
// r = lane & 3: one row, one column, one row), staged through S.t viewed as 8 x 16 ints.  Returns the number of
// non-zero levels of the lane's block; levels go to S.lv[256 + blk * 16 ..], the reconstruction to `dst`.
__device__ __forceinline__ int hv_chroma_cu(HvScratch& S, int lane, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int stride,
                                            uint32_t pred, int qpc, bool intra) {
    const int blk = lane >> 2, r = lane & 3;
    int* T = &S.t[0][0] + blk * 20;     // 16 + padding
    {
        const uint32_t s4 = ld_u32(src);
        int d[4], o[4];
#pragma unroll
        for (int x = 0; x < 4; x++) d[x] = (int)((s4 >> (8 * x)) & 255) - (int)((pred >> (8 * x)) & 255);
        hv_fwd4(d, o);
#pragma unroll
        for (int x = 0; x < 4; x++) T[4 * r + x] = (o[x] + 1) >> 1;        // stage shift log2(4) + 8 - 9 = 1
    }
    __syncwarp();
    int nz = 0;
    {
        const int qbits = 14 + qpc / 6 + 5, qs = hevc_quant_scale[qpc % 6], ls = hevc_level_scale[qpc % 6], sh = qpc / 6;
        const long long offs = (long long)(intra ? 171 : 85) << (qbits - 9);
        int in[4] = {T[r], T[4 + r], T[8 + r], T[12 + r]}, w[4], c[4], g[4];
        hv_fwd4(in, w);
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const int l = hv_quant1((w[y] + 128) >> 8, qs, offs, qbits);
            S.lv[256 + blk * 16 + 4 * y + r] = (int16_t)l;
            nz += l != 0;
            c[y] = hv_dequant1(l, ls, sh, 5);
        }
        hv_inv4(c, g);
#pragma unroll
        for (int y = 0; y < 4; y++) T[4 * y + r] = vcp_clip3(-32768, 32767, (g[y] + 64) >> 7);
    }
    nz += __shfl_xor_sync(0xffffffffu, nz, 1);
    nz += __shfl_xor_sync(0xffffffffu, nz, 2);
    __syncwarp();
    uint32_t outw = pred;
    if (nz) {
        int in[4] = {T[4 * r], T[4 * r + 1], T[4 * r + 2], T[4 * r + 3]}, o[4];
        hv_inv4(in, o);
        outw = 0;
#pragma unroll
        for (int x = 0; x < 4; x++) outw |= (uint32_t)vcp_clip255((int)((pred >> (8 * x)) & 255) + ((o[x] + 2048) >> 12)) << (8 * x);
    }
    *reinterpret_cast<uint32_t*>(dst) = outw;
    __syncwarp();
    return nz;
}

// One 4x4 chroma transform block per lane, in registers.  pred: 4 rows packed; returns nz, writes levels and recon.
__device__ __forceinline__ int hv_chroma_tu(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int stride, const uint32_t pred[4],
                                            int qpc, bool intra, int16_t* lv) {
    int t[16], w[16];
#pragma unroll
    for (int y = 0; y < 4; y++) {
        const uint32_t s4 = ld_u32(src + (size_t)y * stride), p4 = pred[y];
        int d[4], o[4];
#pragma unroll
        for (int x = 0; x < 4; x++) d[x] = (int)((s4 >> (8 * x)) & 255) - (int)((p4 >> (8 * x)) & 255);
        hv_fwd4(d, o);
#pragma unroll
        for (int x = 0; x < 4; x++) t[4 * y + x] = (o[x] + 1) >> 1;       // stage shift log2(4) + 8 - 9 = 1
    }
#pragma unroll
    for (int x = 0; x < 4; x++) {
        int in[4] = {t[x], t[4 + x], t[8 + x], t[12 + x]}, o[4];
        hv_fwd4(in, o);
#pragma unroll
        for (int y = 0; y < 4; y++) w[4 * y + x] = (o[y] + 128) >> 8;     // log2(4) + 6
    }
    const int qbits = 14 + qpc / 6 + 5, qs = hevc_quant_scale[qpc % 6], ls = hevc_level_scale[qpc % 6], sh = qpc / 6;
    const long long offs = (long long)(intra ? 171 : 85) << (qbits - 9);
    int nz = 0, c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int l = hv_quant1(w[i], qs, offs, qbits);
        lv[i] = (int16_t)l;
        nz += l != 0;
        c[i] = hv_dequant1(l, ls, sh, 5);
    }
    if (nz) {
        int g[16];
#pragma unroll
        for (int x = 0; x < 4; x++) {     // columns first
            int in[4] = {c[x], c[4 + x], c[8 + x], c[12 + x]}, o[4];
            hv_inv4(in, o);
#pragma unroll
            for (int y = 0; y < 4; y++) g[4 * y + x] = vcp_clip3(-32768, 32767, (o[y] + 64) >> 7);
        }
#pragma unroll
        for (int y = 0; y < 4; y++) {
            int o[4];
            hv_inv4(&g[4 * y], o);
            uint32_t outw = 0;
#pragma unroll
            for (int x = 0; x < 4; x++) outw |= (uint32_t)vcp_clip255((int)((pred[y] >> (8 * x)) & 255) + ((o[x] + 2048) >> 12)) << (8 * x);
            *reinterpret_cast<uint32_t*>(dst + (size_t)y * stride) = outw;
        }
    } else {
#pragma unroll
        for (int y = 0; y < 4; y++) *reinterpret_cast<uint32_t*>(dst + (size_t)y * stride) = pred[y];
    }
    return nz;
}

// ---- inter CUs -------------------------------------------------------------------------------------
constexpr int HP_WARPS = 4;

__global__ void __launch_bounds__(HP_WARPS * 32) hevc_p_recon_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ HvScratch scr[HP_WARPS];
    __shared__ __align__(16) uint32_t qwin_all[HP_WARPS][HQ_WIN][8];   // quarter-sample vectors: the integer samples -3 .. 20 around the block (24 bytes per row used)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * HP_WARPS + warp;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    HvScratch& S = scr[warp];
    const int n = vcp_frame_of(s, gi);
    const int slot = vcp_rec_slot(s, gi, s.t), rslot = vcp_rec_slot(s, gi, s.t - 1);
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    const int qp = b.qp[n], qpc = hevc_chroma_qp(qp);
    const size_t o = (size_t)gi * g.nmb + mbi;
    if (b.mbtype[o] == VCP_MB_I16) return;   // flagged intra by the refine: coded by hevc_i_fix_kernel once we are done
    const short2 mv = b.mv[o];
    // luma prediction: full-sample vector, a copy of the reference
    {
        const int row = lane >> 1, hx = (lane & 1) * 8;
        const uint8_t* r0 = vcp_rec_luma(b, g, rslot) + g.yoff + (ptrdiff_t)(16 * my + row) * g.ys + 16 * mx + hx;
        if ((mv.x | mv.y) & 1) {
            // quarter-sample vector (hevc_subpel = 2): interpolate from the integer samples (vcp_hevc_qpel.cuh), window
            // origin at the first tap of sample (0, 0)
            if (lane < HQ_WIN) {
                const uint8_t* wr = vcp_rec_luma(b, g, rslot) + g.yoff + (ptrdiff_t)(16 * my + (mv.y >> 2) - 3 + lane) * g.ys + 16 * mx + (mv.x >> 2) - 3;
                const uint2 a = ld8_unaligned(wr), c = ld8_unaligned(wr + 8), e = ld8_unaligned(wr + 16);
                *reinterpret_cast<uint4*>(&qwin_all[warp][lane][0]) = make_uint4(a.x, a.y, c.x, c.y);
                *reinterpret_cast<uint4*>(&qwin_all[warp][lane][4]) = make_uint4(e.x, e.y, 0u, 0u);
            }
            __syncwarp();
            uint32_t* W = reinterpret_cast<uint32_t*>(&S.t[0][0]);
            hq_hpass(&qwin_all[warp][0][0], 8, 0, mv.x & 3, W, lane);
            __syncwarp();
            *reinterpret_cast<uint2*>(&S.pred[row][hx]) = hq_vpass(W, row, lane & 1, mv.y & 3);
        } else
        // half-sample vectors (hevc_subpel) read the 8-tap planes of the reference; full-sample ones the picture itself
        *reinterpret_cast<uint2*>(&S.pred[row][hx]) = g.hevc_subpel ? hpel_fetch8(r0, mv.x, mv.y, g.ys, g.ysize)
                                                                    : ld8_unaligned(r0 + (ptrdiff_t)(mv.y >> 2) * g.ys + (mv.x >> 2));
    }
    // chroma prediction (8.5.3.3.3.2): the luma vector in 1/8 chroma samples (fractions 0 / 4 with full-sample luma
    // vectors, 0 / 2 / 4 / 6 with half-sample ones), 4-tap filters, horizontal then vertical
    {
        const int pl = lane >> 4, row = (lane >> 1) & 7, x0 = (lane & 1) * 4;
        const int ix = mv.x >> 3, iy = mv.y >> 3, fx = mv.x & 7, fy = mv.y & 7;
        const int f0 = hevc_chroma_filter[fx][0], f1 = hevc_chroma_filter[fx][1], f2 = hevc_chroma_filter[fx][2], f3 = hevc_chroma_filter[fx][3];
        const int g0 = hevc_chroma_filter[fy][0], g1 = hevc_chroma_filter[fy][1], g2 = hevc_chroma_filter[fy][2], g3 = hevc_chroma_filter[fy][3];
        const uint8_t* base = (pl ? b.rec_v : b.rec_u) + (size_t)rslot * g.csize + g.coff + (ptrdiff_t)(8 * my + row + iy) * g.cs + 8 * mx + x0 + ix;
        uint32_t outw = 0;
#pragma unroll
        for (int x = 0; x < 4; x++) {
            const uint8_t* p = base + x;
            int v;
            if (!fx && !fy) v = (int)p[0] << 6;
            else if (!fy) v = f0 * p[-1] + f1 * p[0] + f2 * p[1] + f3 * p[2];
            else if (!fx) v = g0 * p[-g.cs] + g1 * p[0] + g2 * p[g.cs] + g3 * p[2 * g.cs];
            else {
                int t[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { const uint8_t* q = p + (ptrdiff_t)(k - 1) * g.cs; t[k] = f0 * q[-1] + f1 * q[0] + f2 * q[1] + f3 * q[2]; }
                v = (g0 * t[0] + g1 * t[1] + g2 * t[2] + g3 * t[3]) >> 6;
            }
            outw |= (uint32_t)vcp_clip255((v + 32) >> 6) << (8 * x);
        }
        *reinterpret_cast<uint32_t*>(&S.cpred[pl][row][x0]) = outw;
    }
    __syncwarp();
    // luma: eight lanes per 8x8 block
    const int k = lane >> 3, q = lane & 7, bx = (k & 1) * 8, by = (k >> 1) * 8;
    const uint8_t* srcy = b.src_y + (size_t)n * g.ysize + g.yoff + (size_t)(16 * my + by) * g.ys + 16 * mx + bx;
    const int nzy = hv_luma_tu(S, true, k, q, bx, by, srcy, g.ys, qp, false);
    const uint32_t ymask = __ballot_sync(0xffffffffu, nzy > 0);
    // chroma: four lanes per 4x4 block, lane = (plane * 4 + z) * 4 + row
    int nzc;
    {
        const int pl = lane >> 4, z = (lane >> 2) & 3, r = lane & 3, cx = (z & 1) * 4, cy = (z >> 1) * 4 + r;
        const size_t co = g.coff + (size_t)(8 * my + cy) * g.cs + 8 * mx + cx;
        const uint8_t* sc = (pl ? b.src_v : b.src_u) + (size_t)n * g.csize + co;
        uint8_t* dc = (pl ? b.rec_v : b.rec_u) + (size_t)slot * g.csize + co;
        nzc = hv_chroma_cu(S, lane, sc, dc, g.cs, *reinterpret_cast<const uint32_t*>(&S.cpred[pl][cy][cx]), qpc, false);
    }
    const uint32_t cmask = __ballot_sync(0xffffffffu, nzc > 0);
    // reconstruction and levels out
    if (lane < 16) {
        const int bx4 = (lane & 3) * 4, by4 = (lane >> 2) * 4;
        uint8_t* dst = vcp_rec_luma(b, g, slot) + g.yoff + (size_t)(16 * my + by4) * g.ys + 16 * mx + bx4;
#pragma unroll
        for (int y = 0; y < 4; y++) *reinterpret_cast<uint32_t*>(dst + (size_t)y * g.ys) = *reinterpret_cast<const uint32_t*>(&S.pred[by4 + y][bx4]);
    }
    for (int i = lane; i < 48; i += 32) reinterpret_cast<uint4*>(b.levels + o * VCP_LV_STRIDE)[i] = reinterpret_cast<const uint4*>(S.lv)[i];
    if (lane == 0) {
        uint32_t cbf_y = 0, cbf_c = 0;
#pragma unroll
        for (int z = 0; z < 4; z++) cbf_y |= ((ymask >> (8 * z)) & 1u) << z;
#pragma unroll
        for (int z = 0; z < 8; z++) cbf_c |= ((cmask >> (4 * z)) & 1u) << z;
        b.cbp[o] = (uint8_t)cbf_y;
        b.modes[o] = (uint8_t)cbf_c;                       // cbf_cb in bits 0-3, cbf_cr in bits 4-7
    }
}

// ---- intra CUs (IDR pictures), wavefront ---------------------------------------------------------------
constexpr int HI_WARPS = 16;

// DC prediction of one transform block from the reconstruction of the current picture (8.4.4.2.2 / 8.4.4.2.5).
// Only the N samples above and the N to the left enter a DC prediction; with 16x16 CTBs in raster order their
// substitution rules reduce to: nothing available -> 128; left missing -> first top sample; top missing ->
// topmost left sample.  Returns the four (luma: eight) rows through `rows`.
__device__ __forceinline__ void hv_dc_pred(const uint8_t* __restrict__ rec, int stride, bool aL, bool aT, int n, bool luma, uint8_t* out /* n*n */) {
    int top[8], left[8];
    if (aT) for (int i = 0; i < n; i++) top[i] = rec[-(ptrdiff_t)stride + i];
    if (aL) for (int i = 0; i < n; i++) left[i] = rec[(ptrdiff_t)i * stride - 1];
    if (!aT && !aL) { for (int i = 0; i < n; i++) top[i] = left[i] = 128; }
    else if (!aL) { for (int i = 0; i < n; i++) left[i] = top[0]; }
    else if (!aT) { for (int i = 0; i < n; i++) top[i] = left[0]; }
    int sum = n;
    for (int i = 0; i < n; i++) sum += top[i] + left[i];
    const int dc = sum >> (n == 8 ? 4 : 3);
    for (int i = 0; i < n * n; i++) out[i] = (uint8_t)dc;
    if (luma) {
        out[0] = (uint8_t)((left[0] + 2 * dc + top[0] + 2) >> 2);
        for (int x = 1; x < n; x++) out[x] = (uint8_t)((top[x] + 3 * dc + 2) >> 2);
        for (int y = 1; y < n; y++) out[y * n] = (uint8_t)((left[y] + 3 * dc + 2) >> 2);
    }
}

__device__ void hv_encode_intra_cu(const VcpGeom& g, const VcpBufs& b, HvScratch& S, int n, int slot, int gi, int mx, int my, int row0,
                                   int qp, int qpc, int lane) {
    const size_t o = (size_t)gi * g.nmb + my * g.mbw + mx;
    uint8_t* ry = vcp_rec_luma(b, g, slot) + g.yoff + (size_t)(16 * my) * g.ys + 16 * mx;
    uint32_t cbf_y = 0, cbf_c = 0;
    for (int z = 0; z < 4; z++) {
        const int bx = (z & 1) * 8, by = (z >> 1) * 8;
        const bool aL = bx > 0 || mx > 0, aT = by > 0 || my > row0;
        // luma prediction by lane 0 into the tile, chroma by lanes 8, 9 into registers
        if (lane == 0) {
            uint8_t tmp[64];
            hv_dc_pred(ry + (size_t)by * g.ys + bx, g.ys, aL, aT, 8, true, tmp);
            for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) S.pred[by + y][bx + x] = tmp[8 * y + x];
        }
        __syncwarp();
        const uint8_t* srcy = b.src_y + (size_t)n * g.ysize + g.yoff + (size_t)(16 * my + by) * g.ys + 16 * mx + bx;
        const int nzy = hv_luma_tu(S, lane < 8, z, lane & 7, bx, by, srcy, g.ys, qp, true);
        if (__shfl_sync(0xffffffffu, nzy, 0) > 0) cbf_y |= 1u << z;
        // this TU's luma reconstruction must be in memory before the next TU predicts from it
        if (lane < 16) {
            const int r = lane >> 1, hx = (lane & 1) * 4;
            *reinterpret_cast<uint32_t*>(ry + (size_t)(by + r) * g.ys + bx + hx) = *reinterpret_cast<const uint32_t*>(&S.pred[by + r][bx + hx]);
        }
        int nzc = 0;
        if (lane == 8 || lane == 9) {
            const int pl = lane - 8, cx = (z & 1) * 4, cy = (z >> 1) * 4;
            const size_t co = g.coff + (size_t)(8 * my + cy) * g.cs + 8 * mx + cx;
            const uint8_t* sc = (pl ? b.src_v : b.src_u) + (size_t)n * g.csize + co;
            uint8_t* dc = (pl ? b.rec_v : b.rec_u) + (size_t)slot * g.csize + co;
            uint8_t tmp[16];
            hv_dc_pred(dc, g.cs, aL, aT, 4, false, tmp);
            uint32_t pw[4];
            for (int y = 0; y < 4; y++) pw[y] = (uint32_t)tmp[4 * y] | ((uint32_t)tmp[4 * y + 1] << 8) | ((uint32_t)tmp[4 * y + 2] << 16) | ((uint32_t)tmp[4 * y + 3] << 24);
            nzc = hv_chroma_tu(sc, dc, g.cs, pw, qpc, true, &S.lv[256 + pl * 64 + z * 16]);
        }
        const uint32_t cm = __ballot_sync(0xffffffffu, nzc > 0);
        if (cm & 0x100u) cbf_c |= 1u << z;
        if (cm & 0x200u) cbf_c |= 16u << z;
        __syncwarp();
    }
    for (int i = lane; i < 48; i += 32) reinterpret_cast<uint4*>(b.levels + o * VCP_LV_STRIDE)[i] = reinterpret_cast<const uint4*>(S.lv)[i];
    if (lane == 0) {
        b.cbp[o] = (uint8_t)cbf_y; b.modes[o] = (uint8_t)cbf_c; b.mbtype[o] = 0;
        b.mv[o] = make_short2(0, 0); b.mvd[o] = make_short2(0, 0);
    }
}

// grid: x = slice, y = GOP.  FIX = false: IDR picture, every CU.  FIX = true: the CUs of a P picture that the
// refine flagged intra (scene cuts), after hevc_p_recon_kernel -- they predict from the reconstruction around
// them whatever its type; only anti-diagonals that hold flagged CUs cost a barrier, a slice without any leaves
// at once.
template <bool FIX>
__global__ void __launch_bounds__(HI_WARPS * 32) hevc_i_recon_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ HvScratch scr[HI_WARPS];
    __shared__ int todo;
    __shared__ uint32_t dmask[32];   // up to 1024 anti-diagonals
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = blockIdx.x, gi = blockIdx.y + s.g0;
    const int r0 = vcp_slice_first_row(sl, g.slices, g.mbh);
    const int r1 = sl + 1 < g.slices ? vcp_slice_first_row(sl + 1, g.slices, g.mbh) : g.mbh;
    const int rows = r1 - r0;
    const int ndiag = g.mbw + rows - 1;
    if (FIX) {
        if (threadIdx.x == 0) { todo = b.icount[(size_t)gi * g.slices + sl]; b.icount[(size_t)gi * g.slices + sl] = 0; }
        if (threadIdx.x < 32) dmask[threadIdx.x] = 0;
        __syncthreads();
        if (!todo) return;
        for (int i = threadIdx.x; i < rows * g.mbw; i += blockDim.x)
            if (b.mbtype[(size_t)gi * g.nmb + r0 * g.mbw + i] == VCP_MB_I16) {
                const int d = i % g.mbw + i / g.mbw;
                atomicOr(&dmask[(d >> 5) & 31], 1u << (d & 31));
            }
        __syncthreads();
    }
    const int n = vcp_frame_of(s, gi);
    const int slot = vcp_rec_slot(s, gi, s.t);
    const int qp = b.qp[n], qpc = hevc_chroma_qp(qp);
    for (int d = 0; d < ndiag; d++) {
        if (FIX && !((dmask[(d >> 5) & 31] >> (d & 31)) & 1)) continue;
        const int k0 = d - (g.mbw - 1) > 0 ? d - (g.mbw - 1) : 0;
        const int k1 = d < rows - 1 ? d : rows - 1;
        for (int k = k0 + warp; k <= k1; k += HI_WARPS) {
            const int mx = d - k, my = r0 + k;
            if (FIX && b.mbtype[(size_t)gi * g.nmb + my * g.mbw + mx] != VCP_MB_I16) continue;
            hv_encode_intra_cu(g, b, scr[warp], n, slot, gi, mx, my, r0, qp, qpc, lane);
        }
        __syncthreads();
    }
}

// ---- merge / AMVP / skip, once every vector of the picture is final -----------------------------------------
__global__ void __launch_bounds__(128) hevc_cuinfo_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    const int mbi = blockIdx.x * blockDim.x + threadIdx.x;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    const size_t base = (size_t)gi * g.nmb;
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    const int row0 = vcp_row_first(b, my);
    // neighbours: 0 A1 left, 1 B1 above, 2 B0 above-right, 3 B2 above-left (A0 is never decoded before us)
    const int nx[4] = {mx - 1, mx, mx + 1, mx - 1}, ny[4] = {my, my - 1, my - 1, my - 1};
    bool av[4], inter[4]; int vx[4], vy[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        av[k] = nx[k] >= 0 && nx[k] < g.mbw && ny[k] >= row0;
        inter[k] = false; vx[k] = vy[k] = 0;
        if (av[k]) {
            const size_t o = base + ny[k] * g.mbw + nx[k];
            if (b.mbtype[o] != 0) { inter[k] = true; const short2 v = b.mv[o]; vx[k] = v.x; vy[k] = v.y; }
        }
    }
    const size_t o = base + mbi;
    if (b.mbtype[o] == 0) return;     // intra
    const short2 mv = b.mv[o];
    // merge candidate 0: A1, B1, B0, B2, else the zero vector
    int mgx = 0, mgy = 0;
    { const int k = inter[0] ? 0 : inter[1] ? 1 : inter[2] ? 2 : inter[3] ? 3 : -1; if (k >= 0) { mgx = vx[k]; mgy = vy[k]; } }
    const int cbf_any = (b.cbp[o] & 15) | b.modes[o];
    if (mgx == mv.x && mgy == mv.y) {
        b.cbp[o] = (uint8_t)((b.cbp[o] & 15) | 16);           // merge_flag
        b.mvd[o] = make_short2(0, 0);
        if (!cbf_any) b.mbtype[o] = 2;                           // cu_skip_flag
        return;
    }
    // AMVP (8.5.3.2.6/7), one reference picture, no temporal candidate
    bool have_a = inter[0], have_b = inter[2] || inter[1] || inter[3];
    int ax = vx[0], ay = vy[0];
    const int kb = inter[2] ? 2 : inter[1] ? 1 : 3;
    const int bx = have_b ? vx[kb] : 0, by = have_b ? vy[kb] : 0;
    if (!av[0] && have_b) { have_a = true; ax = bx; ay = by; }   // isScaledFlag == 0: B stands in for A
    int lx[2] = {0, 0}, ly[2] = {0, 0}, cnt = 0;
    if (have_a) { lx[cnt] = ax; ly[cnt] = ay; cnt++; }
    if (have_b && !(have_a && ax == bx && ay == by)) { lx[cnt] = bx; ly[cnt] = by; cnt++; }
    const int c0 = vcp_se_len(mv.x - lx[0]) + vcp_se_len(mv.y - ly[0]);
    const int c1 = vcp_se_len(mv.x - lx[1]) + vcp_se_len(mv.y - ly[1]);
    const int idx = c1 < c0 ? 1 : 0;
    b.cbp[o] = (uint8_t)((b.cbp[o] & 15) | (idx << 5));
    b.mvd[o] = make_short2((short)(mv.x - lx[idx]), (short)(mv.y - ly[idx]));
}

// ---- in-loop deblocking (8.7.2) ------------------------------------------------------------------------------
// All vertical edges of the picture, then (second launch) all horizontal edges on the result.  Edges sit on the
// 8x8 luma grid and a filter reads 4 / changes at most 3 samples on each side, so within one direction every
// 4-line edge segment is independent: thread = segment, no wavefront (unlike H.264's K4).  The thread holds the
// 8 x 4 samples around its segment in registers (vertical: two words per row; horizontal: one word per row, the
// four lines are the bytes), decides (bS, dE, dEp, dEq) and writes the block back.  Chroma (bS = 2 only, i.e. IDR
// pictures and intra CUs of P pictures; 16-sample luma grid) rides on the same thread.  Slices are not filtered across
// (pps_loop_filter_across_slices_enabled_flag = 0).
template <bool VERTICAL>
__global__ void __launch_bounds__(256) hevc_deblock_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    const int nx = VERTICAL ? g.cw >> 3 : g.cw >> 2, ny = VERTICAL ? g.ch >> 2 : g.ch >> 3;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nx * ny) return;
    const int gi = blockIdx.y + s.g0;
    const int n = vcp_frame_of(s, gi), slot = vcp_rec_slot(s, gi, s.t);
    const int x = (tid % nx) * (VERTICAL ? 8 : 4), y = (tid / nx) * (VERTICAL ? 4 : 8);
    const int xp = VERTICAL ? x - 1 : x, yp = VERTICAL ? y : y - 1;
    if (xp < 0 || yp < 0) return;
    if (!VERTICAL && (y & 15) == 0 && vcp_row_first(b, y >> 4) == (y >> 4)) return;     // top edge of a slice
    // boundary strength
    const size_t base = (size_t)gi * g.nmb;
    const size_t oq = base + (size_t)(y >> 4) * g.mbw + (x >> 4), op = base + (size_t)(yp >> 4) * g.mbw + (xp >> 4);
    int bs;
    if (s.t == 0 || b.mbtype[oq] == VCP_MB_I16 || b.mbtype[op] == VCP_MB_I16) bs = 2;   // intra on either side (IDR: every CU)
    else {
        const int zq = ((y >> 3) & 1) * 2 + ((x >> 3) & 1), zp = ((yp >> 3) & 1) * 2 + ((xp >> 3) & 1);
        if (((b.cbp[oq] >> zq) | (b.cbp[op] >> zp)) & 1) bs = 1;    // a transform block with coefficients on either side
        else if (op == oq) return;
        else {
            const short2 vq = b.mv[oq], vp = b.mv[op];
            if (vcp_iabs(vp.x - vq.x) < 4 && vcp_iabs(vp.y - vq.y) < 4) return;
            bs = 1;
        }
    }
    const int qp = b.qp[n];
    const int beta = hevc_beta_tab[qp], tc = hevc_tc_tab[min(53, qp + 2 * (bs - 1))];
    uint8_t* Y = vcp_rec_luma(b, g, slot) + g.yoff + (size_t)y * g.ys + x;
    // P[i][l], Q[i][l]: distance i from the edge, line l along it
    int P[4][4], Q[4][4];
    if (VERTICAL) {
#pragma unroll
        for (int l = 0; l < 4; l++) {
            const uint32_t wp = ld_u32(Y + (size_t)l * g.ys - 4), wq = ld_u32(Y + (size_t)l * g.ys);
#pragma unroll
            for (int i = 0; i < 4; i++) { P[i][l] = (int)((wp >> (8 * (3 - i))) & 255); Q[i][l] = (int)((wq >> (8 * i)) & 255); }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t wp = ld_u32(Y - (ptrdiff_t)(i + 1) * g.ys), wq = ld_u32(Y + (size_t)i * g.ys);
#pragma unroll
            for (int l = 0; l < 4; l++) { P[i][l] = (int)((wp >> (8 * l)) & 255); Q[i][l] = (int)((wq >> (8 * l)) & 255); }
        }
    }
    bool changed = false;
    if (beta) {
        const int dp0 = vcp_iabs(P[2][0] - 2 * P[1][0] + P[0][0]), dp3 = vcp_iabs(P[2][3] - 2 * P[1][3] + P[0][3]);
        const int dq0 = vcp_iabs(Q[2][0] - 2 * Q[1][0] + Q[0][0]), dq3 = vcp_iabs(Q[2][3] - 2 * Q[1][3] + Q[0][3]);
        const int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3;
        if (dpq0 + dpq3 < beta) {
            const int tc25 = (5 * tc + 1) >> 1;
            const bool s0 = 2 * dpq0 < (beta >> 2) && vcp_iabs(P[3][0] - P[0][0]) + vcp_iabs(Q[0][0] - Q[3][0]) < (beta >> 3) && vcp_iabs(P[0][0] - Q[0][0]) < tc25;
            const bool s3 = 2 * dpq3 < (beta >> 2) && vcp_iabs(P[3][3] - P[0][3]) + vcp_iabs(Q[0][3] - Q[3][3]) < (beta >> 3) && vcp_iabs(P[0][3] - Q[0][3]) < tc25;
            const int side = (beta + (beta >> 1)) >> 3;
            const bool dep = dp0 + dp3 < side, deq = dq0 + dq3 < side;
            changed = true;
#pragma unroll
            for (int l = 0; l < 4; l++) {
                const int p0 = P[0][l], p1 = P[1][l], p2 = P[2][l], p3 = P[3][l], q0 = Q[0][l], q1 = Q[1][l], q2 = Q[2][l], q3 = Q[3][l];
                if (s0 && s3) {
                    P[0][l] = vcp_clip3(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
                    P[1][l] = vcp_clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
                    P[2][l] = vcp_clip3(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
                    Q[0][l] = vcp_clip3(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
                    Q[1][l] = vcp_clip3(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
                    Q[2][l] = vcp_clip3(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
                } else {
                    int d = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
                    if (vcp_iabs(d) < tc * 10) {
                        d = vcp_clip3(-tc, tc, d);
                        P[0][l] = vcp_clip255(p0 + d);
                        Q[0][l] = vcp_clip255(q0 - d);
                        if (dep) P[1][l] = vcp_clip255(p1 + vcp_clip3(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + d) >> 1));
                        if (deq) Q[1][l] = vcp_clip255(q1 + vcp_clip3(-(tc >> 1), tc >> 1, (((q2 + q0 + 1) >> 1) - q1 - d) >> 1));
                    }
                }
            }
        }
    }
    if (changed) {
        if (VERTICAL) {
#pragma unroll
            for (int l = 0; l < 4; l++) {
                uint32_t wp = 0, wq = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) { wp |= (uint32_t)P[i][l] << (8 * (3 - i)); wq |= (uint32_t)Q[i][l] << (8 * i); }
                *reinterpret_cast<uint32_t*>(Y + (size_t)l * g.ys - 4) = wp;
                *reinterpret_cast<uint32_t*>(Y + (size_t)l * g.ys) = wq;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++) {      // P[3], Q[3] never change
                uint32_t wp = 0, wq = 0;
#pragma unroll
                for (int l = 0; l < 4; l++) { wp |= (uint32_t)P[i][l] << (8 * l); wq |= (uint32_t)Q[i][l] << (8 * l); }
                *reinterpret_cast<uint32_t*>(Y - (ptrdiff_t)(i + 1) * g.ys) = wp;
                *reinterpret_cast<uint32_t*>(Y + (size_t)i * g.ys) = wq;
            }
        }
    }
    // chroma: bS 2 on the 8-sample chroma grid, two chroma lines per segment
    if (bs == 2 && ((VERTICAL ? x : y) & 15) == 0) {
        const int tcc = hevc_tc_tab[min(53, hevc_chroma_qp(qp) + 2)];
        if (tcc == 0) return;
        const ptrdiff_t step = VERTICAL ? 1 : g.cs, line = VERTICAL ? g.cs : 1;
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            uint8_t* C = (pl ? b.rec_v : b.rec_u) + (size_t)slot * g.csize + g.coff + (size_t)(y >> 1) * g.cs + (x >> 1);
#pragma unroll
            for (int l = 0; l < 2; l++) {
                uint8_t* c = C + l * line;
                const int p0 = c[-step], p1 = c[-2 * step], q0 = c[0], q1 = c[step];
                const int d = vcp_clip3(-tcc, tcc, (((q0 - p0) << 2) + p1 - q1 + 4) >> 3);
                c[-step] = (uint8_t)vcp_clip255(p0 + d);
                c[0] = (uint8_t)vcp_clip255(q0 - d);
            }
        }
    }
}

// ---- sample adaptive offset (8.7.3), luma edge offsets ----------------------------------------------------------
// Decision and application work on the DEBLOCKED picture D (plane G of the slot).  warp = coding tree block, lane =
// 8 samples of one row with their 3 x 10 neighbourhood in registers.  Statistics of (source - D) per edge class and
// category -> 32 warp reductions -> lane 0 picks "off" or the cheapest class (vcp_algo.h: vcp_sao_class_cost, mirrored
// by the oracle) and stores it in the CTB's record (nnz[0] = 0 / 1 + class, nnz[1..4] = offsets) for the
// binarisation; CTBs that switch SAO on write their filtered samples to the slot's second luma plane (scratch until
// the half-sample planes are built), and hevc_sao_copy_kernel moves them back once every CTB has read its
// neighbours' unfiltered samples.  A sample whose neighbour is outside the picture or in another slice keeps its value.
struct SaoRow { int up[10], mid[10], dn[10]; bool up_ok, dn_ok; };

__device__ __forceinline__ void hv_sao_load(const VcpGeom& g, const VcpBufs& b, const uint8_t* __restrict__ D, int mx, int my, int lane, SaoRow& R) {
    const int r = lane >> 1, x0 = 16 * mx + (lane & 1) * 8, y = 16 * my + r;
    R.up_ok = r > 0 || my > vcp_row_first(b, my);
    R.dn_ok = r < 15 || (my + 1 < g.mbh && vcp_row_slice(b, my + 1) == vcp_row_slice(b, my));
    const int xl = x0 > 0 ? x0 - 1 : 0, xr = x0 + 8 < g.cw ? x0 + 8 : g.cw - 1;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int yy = vcp_clip3(0, g.ch - 1, y + k - 1);
        const uint8_t* p = D + (size_t)yy * g.ys;
        const uint2 w = *reinterpret_cast<const uint2*>(p + x0);
        int* a = k == 0 ? R.up : k == 1 ? R.mid : R.dn;
        a[0] = p[xl]; a[9] = p[xr];
#pragma unroll
        for (int j = 0; j < 8; j++) a[1 + j] = (int)(((j < 4 ? w.x : w.y) >> (8 * (j & 3))) & 255);
    }
}
// edge category (0 = none, 1..4) of sample j (0..7) of the lane's row for class cls
__device__ __forceinline__ int hv_sao_cat(const VcpGeom& g, const SaoRow& R, int x, int j, int cls) {
    const bool lr = x > 0 && x + 1 < g.cw, ud = R.up_ok && R.dn_ok;
    int n0, n1;
    bool ok;
    if (cls == 0) { n0 = R.mid[j]; n1 = R.mid[j + 2]; ok = lr; }
    else if (cls == 1) { n0 = R.up[j + 1]; n1 = R.dn[j + 1]; ok = ud; }
    else if (cls == 2) { n0 = R.up[j]; n1 = R.dn[j + 2]; ok = lr && ud; }
    else { n0 = R.up[j + 2]; n1 = R.dn[j]; ok = lr && ud; }
    if (!ok) return 0;
    const int c = R.mid[j + 1];
    const int sg = (c < n0 ? -1 : c > n0 ? 1 : 0) + (c < n1 ? -1 : c > n1 ? 1 : 0);
    return sg == -2 ? 1 : sg == -1 ? 2 : sg == 1 ? 3 : sg == 2 ? 4 : 0;
}

constexpr int SAO_WARPS = 4;
__global__ void __launch_bounds__(SAO_WARPS * 32) hevc_sao_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * SAO_WARPS + warp;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    const int n = vcp_frame_of(s, gi), slot = vcp_rec_slot(s, gi, s.t);
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    uint8_t* D = vcp_rec_luma(b, g, slot) + g.yoff;
    SaoRow R;
    hv_sao_load(g, b, D, mx, my, lane, R);
    const int r = lane >> 1, x0 = 16 * mx + (lane & 1) * 8, y = 16 * my + r;
    const uint2 sw = *reinterpret_cast<const uint2*>(b.src_y + (size_t)n * g.ysize + g.yoff + (size_t)y * g.ys + x0);
    int sum[4][4], cnt[4][4];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < 4; k++) { sum[c][k] = 0; cnt[c][k] = 0; }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int diff = (int)(((j < 4 ? sw.x : sw.y) >> (8 * (j & 3))) & 255) - R.mid[j + 1];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int k = hv_sao_cat(g, R, x0 + j, j, c);
#pragma unroll
            for (int q = 0; q < 4; q++) { sum[c][q] += k == q + 1 ? diff : 0; cnt[c][q] += k == q + 1; }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < 4; k++) { sum[c][k] = warp_sum(sum[c][k]); cnt[c][k] = warp_sum(cnt[c][k]); }
    // decision (every lane computes the same)
    const int lam = vcp_lambda(b.qp[n]), lam2 = lam * lam;
    long long best = lam2;
    int type = 0, off[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int o[4];
        const long long cost = vcp_sao_class_cost(sum[c], cnt[c], lam2, o);
        if (cost < best && (o[0] | o[1] | o[2] | o[3])) { best = cost; type = 1 + c; off[0] = o[0]; off[1] = o[1]; off[2] = o[2]; off[3] = o[3]; }
    }
    uint8_t* rec = b.nnz + ((size_t)gi * g.nmb + mbi) * 24;
    if (lane == 0) rec[0] = (uint8_t)type;
    if (lane >= 1 && lane <= 4) rec[lane] = (uint8_t)(int8_t)off[lane - 1];
    if (!type) return;
    // filtered samples of this CTB -> scratch plane
    uint32_t o4[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int k = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) if (type == 1 + c) k = hv_sao_cat(g, R, x0 + j, j, c);
        const int ov = k == 1 ? off[0] : k == 2 ? off[1] : k == 3 ? off[2] : k == 4 ? off[3] : 0;
        const int v = vcp_clip255(R.mid[j + 1] + ov);
        o4[j >> 2] |= (uint32_t)v << (8 * (j & 3));
    }
    *reinterpret_cast<uint2*>(D + g.ysize + (size_t)y * g.ys + x0) = make_uint2(o4[0], o4[1]);
}

__global__ void __launch_bounds__(SAO_WARPS * 32) hevc_sao_copy_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * SAO_WARPS + warp;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    if (!b.nnz[((size_t)gi * g.nmb + mbi) * 24]) return;
    const int slot = vcp_rec_slot(s, gi, s.t);
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    uint8_t* D = vcp_rec_luma(b, g, slot) + g.yoff + (size_t)(16 * my + (lane >> 1)) * g.ys + 16 * mx + (lane & 1) * 8;
    *reinterpret_cast<uint2*>(D) = *reinterpret_cast<const uint2*>(D + g.ysize);
}

}  // namespace

void vcp_launch_hevc_p_recon(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + HP_WARPS - 1) / HP_WARPS, s.ngop);
    hevc_p_recon_kernel<<<grid, HP_WARPS * 32, 0, st>>>(g, b, s);
}
void vcp_launch_hevc_i_recon(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid(g.slices, s.ngop);
    hevc_i_recon_kernel<false><<<grid, HI_WARPS * 32, 0, st>>>(g, b, s);
}
void vcp_launch_hevc_i_fix(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid(g.slices, s.ngop);
    hevc_i_recon_kernel<true><<<grid, HI_WARPS * 32, 0, st>>>(g, b, s);
}
void vcp_launch_hevc_cuinfo(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + 127) / 128, s.ngop);
    hevc_cuinfo_kernel<<<grid, 128, 0, st>>>(g, b, s);
}
void vcp_launch_hevc_deblock(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    const int nv = (g.cw >> 3) * (g.ch >> 2), nh = (g.cw >> 2) * (g.ch >> 3);
    hevc_deblock_kernel<true><<<dim3((nv + 255) / 256, s.ngop), 256, 0, st>>>(g, b, s);
    hevc_deblock_kernel<false><<<dim3((nh + 255) / 256, s.ngop), 256, 0, st>>>(g, b, s);
}
void vcp_launch_hevc_sao(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + SAO_WARPS - 1) / SAO_WARPS, s.ngop);
    hevc_sao_kernel<<<grid, SAO_WARPS * 32, 0, st>>>(g, b, s);
}
void vcp_launch_hevc_sao_copy(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + SAO_WARPS - 1) / SAO_WARPS, s.ngop);
    hevc_sao_copy_kernel<<<grid, SAO_WARPS * 32, 0, st>>>(g, b, s);
}
