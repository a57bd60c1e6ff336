// K5 (CABAC) — H.264 9.3: binarisation + context modelling (macroblock-parallel) and the
// arithmetic coder (one sequential coder per slice, every picture of the batch in parallel).
//
// CABAC adapts its probability states along the whole slice, so the arithmetic coder is serial
// per slice by construction.  What is NOT serial is everything in front of it: which bins a
// macroblock produces, and in which contexts, depends only on its own record and on its
// left/top neighbours' records.  The stage is therefore split:
//
//   cabac_bins_kernel   : per picture step, one warp per macroblock, one lane per syntax group
//                         (header, Intra16x16 DC, 16 luma blocks, 2 chroma DC, 8 chroma AC, end_of_slice)
//                         -> a stream of 16-bit bins {ctxIdx, value, kind} in a bump-allocated
//                         arena + a (offset,count) descriptor per macroblock.  Runs on the entropy
//                         side stream next to the reconstruction chain.
//   cabac_encode_kernel : for a batch of pictures of every resident GOP, one WARP per slice: lane 0
//                         walks the bins through the arithmetic coder (byte-wise low/queue/
//                         outstanding form, bit-identical to 9.3.4.2's PutBit procedure) out of
//                         shared memory while the other lanes stage bins and drain bytes with
//                         coalesced accesses.  The chain is latency-bound; slices interleave per SM.
//   cabac_pack_kernel   : NAL encapsulation of those RBSPs (shared with CAVLC, vcp_entropy.cuh).
//
// Replaces x264's cabac.c inside the ffmpeg child (/root/reference/cmd/consumer.go:376-382; the
// h264-cpu preset is High/CABAC by default).  Output bytes are identical to oracle/h264_oracle.c
// (write_slice_data_cabac, cabac_block, cabac_mvd, cabac_encode/bypass/terminate).
#include <cstdlib>

#include "vcp_dev.cuh"

#define VCP_TAB static __device__ const
#include "h264_cabac_tables.h"
#include "h264_tables.h"
#include "hevc_tables.h"
#include "vcp_entropy.cuh"

namespace {

constexpr int CB_WARPS = 4;
constexpr uint32_t BIN_BYPASS = 1u << 11, BIN_TERM = 2u << 11;
constexpr int NCTX = 460;   // ctxIdx 0..459 cover every syntax element of frame-coded 4:2:0 slices

// ---- bin sink: counts, or writes 16-bit bins ----------------------------------------------------
template <bool WRITE>
struct BinSink {
    uint16_t* dst;
    int n;
    __device__ __forceinline__ void put(int ctx, int bin) { if (WRITE) dst[n] = (uint16_t)(ctx | (bin ? 1 << 10 : 0)); n++; }
    __device__ __forceinline__ void bypass(int bin) { if (WRITE) dst[n] = (uint16_t)(BIN_BYPASS | (bin ? 1 << 10 : 0)); n++; }
    __device__ __forceinline__ void term(int bin) { if (WRITE) dst[n] = (uint16_t)(BIN_TERM | (bin ? 1 << 10 : 0)); n++; }
    __device__ __forceinline__ void ueg(uint32_t v, int k) {   // Exp-Golomb order k in bypass bins
        while (v >= (1u << k)) { bypass(1); v -= 1u << k; k++; }
        bypass(0);
        while (k--) bypass((int)((v >> k) & 1));
    }
};

// residual_block_cabac (7.3.5.3.3): c[0..n-1] in scan order (shared memory)
template <bool WRITE>
__device__ __forceinline__ void cabac_block(BinSink<WRITE>& bs, const int16_t* c, int n, int cat, int cbf_inc) {
    // ctxBlockCatOffset: coded_block_flag 0,4,8,12,16 ; sig/last 0,15,29,44,47 ; abs 0,10,20,30,39
    const int cbf_off = 4 * cat;
    const int sig_off = (int)((0x2F2C1D0F00ull >> (8 * cat)) & 255);
    const int abs_off = (int)((0x271E140A00ull >> (8 * cat)) & 255);
    uint32_t mask = 0;
    for (int i = 0; i < n; i++) mask |= (c[i] != 0 ? 1u : 0u) << i;
    bs.put(85 + cbf_off + cbf_inc, mask != 0);
    if (!mask) return;
    const int last = 31 - __clz(mask);
    for (int i = 0; i < n - 1; i++) {
        const int inc = cat == 3 ? (i < 2 ? i : 2) : i;
        const int sig = (mask >> i) & 1;
        bs.put(105 + sig_off + inc, sig);
        if (sig) {
            bs.put(166 + sig_off + inc, i == last);
            if (i == last) break;
        }
    }
    int gt1 = 0, eq1 = 0;
    uint32_t m = mask;
    while (m) {
        const int i = 31 - __clz(m);
        m &= ~(1u << i);
        const int v = c[i];
        const int a = vcp_iabs(v) - 1;
        const int inc = gt1 ? 0 : (1 + eq1 < 4 ? 1 + eq1 : 4);
        bs.put(227 + abs_off + inc, a > 0);
        if (a > 0) {
            const int lim = 4 - (cat == 3);
            const int ctx = 227 + abs_off + 5 + (gt1 < lim ? gt1 : lim);
            const int ones = a < 14 ? a : 14;
            for (int k = 1; k < ones; k++) bs.put(ctx, 1);
            if (a < 14) bs.put(ctx, 0); else bs.ueg((uint32_t)(a - 14), 0);
            gt1++;
        } else eq1++;
        bs.bypass(v < 0);
    }
}

// ctxBlockCat 5: one 8x8 luma block, 64 scan positions, no coded_block_flag (inferred from the coded
// block pattern); significance contexts by position (table 9-43), levels from ctxIdx 426
template <bool WRITE>
__device__ __forceinline__ void cabac_block8x8(BinSink<WRITE>& bs, const int16_t* c) {
    unsigned long long mask = 0;
    for (int i = 0; i < 64; i++) mask |= (unsigned long long)(c[i] != 0) << i;
    if (!mask) return;
    const int last = 63 - __clzll((long long)mask);
    for (int i = 0; i < 63; i++) {
        const int sig = (int)((mask >> i) & 1);
        bs.put(402 + vcp_cabac_sig8x8[i], sig);
        if (sig) {
            bs.put(417 + vcp_cabac_last8x8[i], i == last);
            if (i == last) break;
        }
    }
    int gt1 = 0, eq1 = 0;
    unsigned long long m = mask;
    while (m) {
        const int i = 63 - __clzll((long long)m);
        m &= ~(1ull << i);
        const int v = c[i];
        const int a = vcp_iabs(v) - 1;
        const int inc = gt1 ? 0 : (1 + eq1 < 4 ? 1 + eq1 : 4);
        bs.put(426 + inc, a > 0);
        if (a > 0) {
            const int ctx = 426 + 5 + (gt1 < 4 ? gt1 : 4);
            const int ones = a < 14 ? a : 14;
            for (int k = 1; k < ones; k++) bs.put(ctx, 1);
            if (a < 14) bs.put(ctx, 0); else bs.ueg((uint32_t)(a - 14), 0);
            gt1++;
        } else eq1++;
        bs.bypass(v < 0);
    }
}

template <bool WRITE>
__device__ __forceinline__ void cabac_mvd(BinSink<WRITE>& bs, int base, int v, int amvd) {
    const int a = vcp_iabs(v);
    bs.put(base + (amvd < 3 ? 0 : amvd > 32 ? 2 : 1), a > 0);
    if (!a) return;
    const int ones = a < 9 ? a : 9;
    for (int k = 1; k < ones; k++) bs.put(base + 3 + (k - 1 < 3 ? k - 1 : 3), 1);
    if (a < 9) bs.put(base + 3 + (a - 1 < 3 ? a - 1 : 3), 0); else bs.ueg((uint32_t)(a - 9), 3);
    bs.bypass(v < 0);
}

struct __align__(16) CbScratch {
    int16_t lv[VCP_LV_STRIDE];
    uint8_t nnz[3][24];   // cur, left, top
};

struct MbCtx {
    int type, cbp, modes;           // this macroblock
    int tA, cbpA, modesA, tB, cbpB, modesB;   // neighbours (t = -1: unavailable)
    short2 mvd, mvdA, mvdB;
    bool idr, last_in_slice, t8x8_mode;
};

// all bins of one macroblock; `lane` selects the syntax group
template <bool WRITE>
__device__ __forceinline__ void mb_bins(BinSink<WRITE>& bs, const CbScratch& S, const MbCtx& M, int lane) {
    const bool aA = M.tA >= 0, aB = M.tB >= 0;
    const bool intra = M.type == VCP_MB_I16;
    const int un = intra ? 1 : 0;
    const int cbpl = M.cbp & 15, cbpc = M.cbp >> 4;
    if (M.type == VCP_MB_PSKIP) {
        if (lane == 0) bs.put(11 + (aA && M.tA != VCP_MB_PSKIP) + (aB && M.tB != VCP_MB_PSKIP), 1);
        if (lane == 28) bs.term(M.last_in_slice);
        return;
    }
    if (lane == 0) {
        if (!M.idr) bs.put(11 + (aA && M.tA != VCP_MB_PSKIP) + (aB && M.tB != VCP_MB_PSKIP), 0);
        if (intra) {
            // mb_type (9.3.2.5): prefix, terminate(0) = not I_PCM, cbp luma, cbp chroma (1 or 2 bins), pred mode (2 bins)
            const int isl = M.idr ? 1 : 0, base = M.idr ? 5 : 17;
            if (!M.idr) bs.put(14, 1);
            bs.put(M.idr ? 3 + aA + aB : 17, 1);
            bs.term(0);
            bs.put(base + 1, cbpl != 0);
            bs.put(base + 2, cbpc != 0);
            if (cbpc) bs.put(base + 2 + isl, cbpc == 2);
            bs.put(base + 3 + isl, (M.modes >> 1) & 1);
            bs.put(base + 3 + 2 * isl, M.modes & 1);
            // intra_chroma_pred_mode
            const int cm = (M.modes >> 2) & 3;
            const int inc = (aA && M.tA == VCP_MB_I16 && ((M.modesA >> 2) & 3)) + (aB && M.tB == VCP_MB_I16 && ((M.modesB >> 2) & 3));
            bs.put(64 + inc, cm > 0);
            if (cm > 0) { bs.put(67, cm > 1); if (cm > 1) bs.put(67, cm > 2); }
        } else {
            bs.put(14, 0); bs.put(15, 0); bs.put(16, 0);   // P_L0_16x16
            const int ax = (aA && M.tA == VCP_MB_P16 ? vcp_iabs(M.mvdA.x) : 0) + (aB && M.tB == VCP_MB_P16 ? vcp_iabs(M.mvdB.x) : 0);
            const int ay = (aA && M.tA == VCP_MB_P16 ? vcp_iabs(M.mvdA.y) : 0) + (aB && M.tB == VCP_MB_P16 ? vcp_iabs(M.mvdB.y) : 0);
            cabac_mvd<WRITE>(bs, 40, M.mvd.x, ax);
            cabac_mvd<WRITE>(bs, 47, M.mvd.y, ay);
            // coded_block_pattern
            const int cA = aA ? (M.tA == VCP_MB_PSKIP ? 0 : M.cbpA) : 0x0f, cB = aB ? (M.tB == VCP_MB_PSKIP ? 0 : M.cbpB) : 0x0f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int a = (k & 1) ? (M.cbp >> (k - 1)) & 1 : (cA >> (k + 1)) & 1;
                const int bb = (k & 2) ? (M.cbp >> (k - 2)) & 1 : (cB >> (k + 2)) & 1;
                bs.put(73 + !a + 2 * !bb, (M.cbp >> k) & 1);
            }
            const int ca = aA && M.tA != VCP_MB_PSKIP ? M.cbpA >> 4 : 0, cb = aB && M.tB != VCP_MB_PSKIP ? M.cbpB >> 4 : 0;
            bs.put(77 + (ca > 0) + 2 * (cb > 0), cbpc > 0);
            if (cbpc) bs.put(77 + 4 + (ca == 2) + 2 * (cb == 2), cbpc == 2);
            if (M.t8x8_mode && cbpl)   // transform_size_8x8_flag
                bs.put(399 + (aA && (M.modesA & 0x80) && M.tA == VCP_MB_P16) + (aB && (M.modesB & 0x80) && M.tB == VCP_MB_P16), (M.modes >> 7) & 1);
        }
        if (intra || M.cbp) bs.put(60, 0);   // mb_qp_delta == 0
    } else if (lane == 1) {
        if (intra) {
            const int fa = aA ? (M.tA == VCP_MB_I16 ? (M.modesA >> 4) & 1 : 0) : un;
            const int fb = aB ? (M.tB == VCP_MB_I16 ? (M.modesB >> 4) & 1 : 0) : un;
            cabac_block<WRITE>(bs, S.lv + VCP_LV_LUMA_DC, 16, 0, fa + 2 * fb);
        }
    } else if (lane < 18) {
        const int blk = lane - 2;
        if (!intra && (M.modes & 0x80)) {
            if (!(blk & 3) && (cbpl & (1 << (blk >> 2)))) cabac_block8x8<WRITE>(bs, S.lv + VCP_LV_LUMA + (blk >> 2) * 64);
        } else if (cbpl & (1 << (blk >> 2))) {
            const int bx = (blk & 1) | ((blk >> 1) & 2), by = ((blk >> 1) & 1) | ((blk >> 2) & 2);
            const int fa = bx > 0 ? S.nnz[0][by * 4 + bx - 1] != 0 : aA ? S.nnz[1][by * 4 + 3] != 0 : un;
            const int fb = by > 0 ? S.nnz[0][(by - 1) * 4 + bx] != 0 : aB ? S.nnz[2][12 + bx] != 0 : un;
            const int16_t* lv = S.lv + VCP_LV_LUMA + blk * 16;
            if (intra) cabac_block<WRITE>(bs, lv + 1, 15, 1, fa + 2 * fb); else cabac_block<WRITE>(bs, lv, 16, 2, fa + 2 * fb);
        }
    } else if (lane < 20) {
        if (cbpc) {
            const int pl = lane - 18;
            const int fa = aA ? (M.tA != VCP_MB_PSKIP && (M.cbpA >> 4) ? (M.modesA >> (5 + pl)) & 1 : 0) : un;
            const int fb = aB ? (M.tB != VCP_MB_PSKIP && (M.cbpB >> 4) ? (M.modesB >> (5 + pl)) & 1 : 0) : un;
            cabac_block<WRITE>(bs, S.lv + VCP_LV_CHROMA_DC + pl * 4, 4, 3, fa + 2 * fb);
        }
    } else if (lane < 28) {
        if (cbpc & 2) {
            const int pl = (lane - 20) >> 2, blk = lane & 3, bx = blk & 1, by = blk >> 1, o = 16 + pl * 4;
            const int fa = bx > 0 ? S.nnz[0][o + by * 2] != 0 : aA ? S.nnz[1][o + by * 2 + 1] != 0 : un;
            const int fb = by > 0 ? S.nnz[0][o + bx] != 0 : aB ? S.nnz[2][o + 2 + bx] != 0 : un;
            cabac_block<WRITE>(bs, S.lv + VCP_LV_CHROMA_AC + (pl * 4 + blk) * 16 + 1, 15, 4, fa + 2 * fb);
        }
    } else if (lane == 28) {
        bs.term(M.last_in_slice);   // end_of_slice_flag
    }
}

// grid: x = macroblock groups, y = GOP of the group
__global__ void __launch_bounds__(CB_WARPS * 32) cabac_bins_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ CbScratch scr[CB_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * CB_WARPS + warp;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    const int n = vcp_frame_of(s, gi);
    const size_t o = (size_t)gi * g.nmb + mbi;
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    const int sl = vcp_row_slice(b, my), row0 = vcp_row_first(b, my);
    const bool aL = mx > 0, aT = my > row0;
    MbCtx M;
    M.type = b.mbtype[o]; M.cbp = b.cbp[o]; M.modes = b.modes[o]; M.mvd = b.mvd[o];
    M.tA = aL ? b.mbtype[o - 1] : -1; M.cbpA = aL ? b.cbp[o - 1] : 0; M.modesA = aL ? b.modes[o - 1] : 0;
    M.tB = aT ? b.mbtype[o - g.mbw] : -1; M.cbpB = aT ? b.cbp[o - g.mbw] : 0; M.modesB = aT ? b.modes[o - g.mbw] : 0;
    M.mvdA = aL ? b.mvd[o - 1] : make_short2(0, 0);
    M.mvdB = aT ? b.mvd[o - g.mbw] : make_short2(0, 0);
    M.idr = s.t == 0;
    M.t8x8_mode = g.t8x8 != 0;
    {
        const int r1 = sl + 1 < g.slices ? vcp_slice_first_row(sl + 1, g.slices, g.mbh) : g.mbh;
        M.last_in_slice = (my == r1 - 1) && (mx == g.mbw - 1);
    }
    CbScratch& S = scr[warp];
    if (M.type != VCP_MB_PSKIP) {
        for (int i = lane; i < VCP_LV_STRIDE * 2 / 16; i += 32)
            reinterpret_cast<uint4*>(S.lv)[i] = reinterpret_cast<const uint4*>(b.levels + o * VCP_LV_STRIDE)[i];
        if (lane < 18) {
            const int w = lane / 6, c = lane % 6;
            const size_t src = w == 0 ? o : (w == 1 ? o - 1 : o - g.mbw);
            const bool ok = w == 0 || (w == 1 ? aL : aT);
            reinterpret_cast<uint32_t*>(S.nnz[w])[c] = ok ? reinterpret_cast<const uint32_t*>(b.nnz + src * 24)[c] : 0u;
        }
    }
    __syncwarp();
    BinSink<false> cnt{nullptr, 0};
    mb_bins<false>(cnt, S, M, lane);
    int incl = cnt.n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long off = 0;
    if (lane == 0) {
        off = atomicAdd(b.bins_cursor, (unsigned long long)total);
        if (off + (unsigned long long)total > b.bins_cap) { atomicExch(b.error_flag, 3); off = ~0ull; }
        else {
            b.mbdesc[(size_t)n * g.nmb + mbi] = make_uint2((uint32_t)off, (uint32_t)total | ((uint32_t)(off >> 32) << 20));
            atomicAdd(&b.slice_bins[(size_t)n * g.slices + sl], (uint32_t)total);
            // rate control sees an estimate of the final bits (vcp_algo.h)
        }
    }
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off == ~0ull) return;
    BinSink<true> wr{b.bins + off + (incl - cnt.n), 0};
    mb_bins<true>(wr, S, M, lane);
}


// ================================= HEVC (k6_hevc.cu records) =======================================
// residual_coding (7.3.8.11) of one transform block: 8x8 luma (four 4x4 sub-blocks) or 4x4 chroma, diagonal
// scan, no sign hiding, no transform skip.  lv: levels in raster order (shared memory).
// Mirrors oracle/hevc_oracle.inc.c hevc_residual bin for bin.
template <bool WRITE>
__device__ __forceinline__ void hevc_residual_bins(BinSink<WRITE>& bs, const int16_t* lv, int log2n, int cidx) {
    const int n = 1 << log2n, nsb = n == 8 ? 4 : 1;
    unsigned long long mask = 0;       // bit 16 i + p: coefficient p of sub-block i (scan order) is non-zero
    for (int i = 0; i < nsb; i++)
        for (int p = 0; p < 16; p++) {
            const int x = (n == 8 ? 4 * hevc_diag2_x[i] : 0) + hevc_diag4_x[p], y = (n == 8 ? 4 * hevc_diag2_y[i] : 0) + hevc_diag4_y[p];
            mask |= (unsigned long long)(lv[y * n + x] != 0) << (16 * i + p);
        }
    if (!mask) return;
    const int last = 63 - __clzll((long long)mask), lastsb = last >> 4;
    {   // last_sig_coeff_{x,y}_prefix (context coded, truncated unary), then the suffixes (bypass)
        const int lx = (n == 8 ? 4 * hevc_diag2_x[lastsb] : 0) + hevc_diag4_x[last & 15];
        const int ly = (n == 8 ? 4 * hevc_diag2_y[lastsb] : 0) + hevc_diag4_y[last & 15];
        int off, shift;
        if (cidx == 0) { off = 3 * (log2n - 2) + ((log2n - 1) >> 2); shift = (log2n + 1) >> 2; }
        else { off = 15; shift = log2n - 2; }
        const int cmax = (log2n << 1) - 1;
#pragma unroll
        for (int comp = 0; comp < 2; comp++) {
            const int v = comp ? ly : lx, base = (comp ? HC_LAST_Y : HC_LAST_X) + off;
            const int prefix = v < 4 ? v : (v < 6 ? 4 : 5);
            for (int i = 0; i < prefix; i++) bs.put(base + (i >> shift), 1);
            if (prefix < cmax) bs.put(base + (prefix >> shift), 0);
        }
        if (lx >= 4) bs.bypass(lx & 1);
        if (ly >= 4) bs.bypass(ly & 1);
    }
    uint32_t csbf = 0;                 // bit ys * 2 + xs
    bool prev_gt1_zero = false, first_sb = true;
    for (int i = lastsb; i >= 0; i--) {
        const int xs = n == 8 ? hevc_diag2_x[i] : 0, ys = n == 8 ? hevc_diag2_y[i] : 0;
        const uint32_t sm = (uint32_t)(mask >> (16 * i)) & 0xffffu;
        bool any = sm != 0;
        const int right = (n == 8 && xs + 1 < 2) ? (csbf >> (ys * 2 + xs + 1)) & 1 : 0;
        const int below = (n == 8 && ys + 1 < 2) ? (csbf >> ((ys + 1) * 2 + xs)) & 1 : 0;
        bool infer_dc = false;
        if (i < lastsb && i > 0) { bs.put(HC_CSBF + (cidx ? 2 : 0) + ((right | below) ? 1 : 0), any); infer_dc = true; }
        else any = true;               // first and last sub-block: coded_sub_block_flag inferred 1
        csbf |= (any ? 1u : 0u) << (ys * 2 + xs);
        if (!any) continue;
        // sig_coeff_flag
        const int start = i == lastsb ? (last & 15) - 1 : 15;
        for (int p = start; p >= 0; p--) {
            if (p == 0 && infer_dc && !(sm & 0xfffeu)) break;    // the only coefficient of a coded sub-block: inferred
            const int xp = hevc_diag4_x[p], yp = hevc_diag4_y[p];
            int sc;
            if (log2n == 2) sc = hevc_sig_ctx_map4[(yp << 2) + xp];
            else if (i == 0 && p == 0) sc = 0;
            else {
                const int pat = right | (below << 1);
                if (pat == 0) sc = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
                else if (pat == 1) sc = yp == 0 ? 2 : yp == 1 ? 1 : 0;
                else if (pat == 2) sc = xp == 0 ? 2 : xp == 1 ? 1 : 0;
                else sc = 2;
                if (cidx == 0) { if (i > 0) sc += 3; sc += 9; } else sc += 9;
            }
            bs.put(HC_SIG + (cidx == 0 ? sc : 27 + sc), (sm >> p) & 1);
        }
        // levels from the highest scan position down
        const int x0 = n == 8 ? 4 * xs : 0, y0 = n == 8 ? 4 * ys : 0;
#define HV_COEF(p_) ((int)lv[(y0 + hevc_diag4_y[p_]) * n + x0 + hevc_diag4_x[p_]])
        int ctxset = (i == 0 || cidx > 0) ? 0 : 2;
        if (!first_sb && prev_gt1_zero) ctxset++;
        first_sb = false;
        int g1ctx = 1, first_g2 = -1, k = 0;
        uint32_t g1mask = 0;
        for (uint32_t m = sm; m && k < 8; k++) {
            const int p = 31 - __clz(m);
            m &= ~(1u << p);
            const int g1 = vcp_iabs(HV_COEF(p)) > 1;
            g1mask |= (uint32_t)g1 << k;
            bs.put(HC_GT1 + (cidx ? 16 : 0) + ctxset * 4 + g1ctx, g1);
            if (g1) { g1ctx = 0; if (first_g2 < 0) first_g2 = k; }
            else if (g1ctx > 0 && g1ctx < 3) g1ctx++;
        }
        prev_gt1_zero = g1ctx == 0;
        int g2flag = 0;
        if (first_g2 >= 0) {
            uint32_t m = sm;
            for (int q = 0; q < first_g2; q++) m &= ~(1u << (31 - __clz(m)));
            g2flag = vcp_iabs(HV_COEF(31 - __clz(m))) > 2;
            bs.put(HC_GT2 + (cidx ? 4 : 0) + ctxset, g2flag);
        }
        for (uint32_t m = sm; m;) { const int p = 31 - __clz(m); m &= ~(1u << p); bs.bypass(HV_COEF(p) < 0); }
        int rice = 0;
        k = 0;
        for (uint32_t m = sm; m; k++) {
            const int p = 31 - __clz(m);
            m &= ~(1u << p);
            const int a = vcp_iabs(HV_COEF(p));
            const int base = k < 8 ? 1 + (int)((g1mask >> k) & 1) + (k == first_g2 ? g2flag : 0) : 1;
            const int thresh = k < 8 ? (k == first_g2 ? 3 : 2) : 1;
            if (base != thresh) continue;
            const int rem = a - base;      // coeff_abs_level_remaining (9.3.3.11)
            if (rem < (3 << rice)) {
                const int len = rem >> rice;
                for (int q = 0; q < len; q++) bs.bypass(1);
                bs.bypass(0);
                for (int q = rice - 1; q >= 0; q--) bs.bypass((rem >> q) & 1);
            } else {
                int len = rice, v = rem - (3 << rice);
                while (v >= (1 << len)) { v -= 1 << len; len++; }
                for (int q = 0; q < 3 + len - rice; q++) bs.bypass(1);
                bs.bypass(0);
                for (int q = len - 1; q >= 0; q--) bs.bypass((v >> q) & 1);
            }
            if (a > 3 * (1 << rice) && rice < 4) rice++;
        }
#undef HV_COEF
    }
}

struct CuCtx {
    int type, cbf_y, cbf_c, merge, mvp_idx;   // cbf_c: Cb in bits 0-3, Cr in bits 4-7
    int tA, tB;                               // neighbours' types (-1: unavailable)
    short2 mvd;
    bool idr, last_in_slice;
    int sao;                                  // -1: SAO not in use; else 0 off / 1 + edge class
    int sao_off[4];
    bool has_left, has_up;                    // CTBs of the same slice to the left / above (sao_merge flags are sent)
};

// all bins of one coding unit (oracle: hevc_write_slice_data / hevc_write_tu_tree); `lane` selects the syntax group:
// 0 = CU header + chroma cbf at depth 0; per transform unit z: 1+4z flags, 2+4z luma, 3+4z Cb, 4+4z Cr; 17 = end_of_slice_segment_flag
template <bool WRITE>
__device__ __forceinline__ void hevc_cu_bins(BinSink<WRITE>& bs, const int16_t* lv, const CuCtx& M, int lane) {
    const bool skip = M.type == 2, intra = M.type == 0;
    const int any_cb = (M.cbf_c & 15) != 0, any_cr = (M.cbf_c >> 4) != 0;
    const bool any = M.cbf_y || M.cbf_c;
    const bool tree = !skip && (intra || any);
    if (lane == 0) {
        if (M.sao >= 0) {   // sao() of the CTU (7.3.8.3): no merging, luma edge offsets only
            if (M.has_left) bs.put(HC_SAO_MERGE, 0);
            if (M.has_up) bs.put(HC_SAO_MERGE, 0);
            bs.put(HC_SAO_TYPE, M.sao != 0);
            if (M.sao) {
                bs.bypass(1);                                    // sao_type_idx_luma = 2
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int a = vcp_iabs(M.sao_off[k]);
                    for (int j = 0; j < a; j++) bs.bypass(1);
                    if (a < 7) bs.bypass(0);
                }
                bs.bypass(((M.sao - 1) >> 1) & 1); bs.bypass((M.sao - 1) & 1);   // sao_eo_class_luma
            }
        }
        if (!M.idr) bs.put(HC_SKIP + (M.tA == 2) + (M.tB == 2), skip);
        if (skip) return;
        if (!M.idr) bs.put(HC_PRED_MODE, intra);
        bs.put(HC_PART_MODE, 1);                                 // PART_2Nx2N
        if (intra) {
            bs.put(HC_PREV_INTRA, 1);                            // prev_intra_luma_pred_flag
            bs.bypass(1); bs.bypass(0);                          // mpm_idx 1: DC
            bs.put(HC_CHROMA_MODE, 0);                           // intra_chroma_pred_mode 4
        } else {
            bs.put(HC_MERGE_FLAG, M.merge);
            if (!M.merge) {
                const int dx = M.mvd.x, dy = M.mvd.y, ax = vcp_iabs(dx), ay = vcp_iabs(dy);
                bs.put(HC_MVD_GT0, ax > 0);
                bs.put(HC_MVD_GT0, ay > 0);
                if (ax) bs.put(HC_MVD_GT1, ax > 1);
                if (ay) bs.put(HC_MVD_GT1, ay > 1);
                if (ax) { if (ax > 1) bs.ueg((uint32_t)(ax - 2), 1); bs.bypass(dx < 0); }
                if (ay) { if (ay > 1) bs.ueg((uint32_t)(ay - 2), 1); bs.bypass(dy < 0); }
                bs.put(HC_MVP_FLAG, M.mvp_idx);
                bs.put(HC_RQT_ROOT_CBF, any);
            }
        }
        if (tree) { bs.put(HC_CBF_CHROMA + 0, any_cb); bs.put(HC_CBF_CHROMA + 0, any_cr); }
    } else if (lane < 17) {
        if (!tree) return;
        const int z = (lane - 1) >> 2, part = (lane - 1) & 3;
        if (part == 0) {
            if (any_cb) bs.put(HC_CBF_CHROMA + 1, (M.cbf_c >> z) & 1);
            if (any_cr) bs.put(HC_CBF_CHROMA + 1, (M.cbf_c >> (4 + z)) & 1);
            bs.put(HC_CBF_LUMA + 0, (M.cbf_y >> z) & 1);
        } else if (part == 1) {
            if ((M.cbf_y >> z) & 1) hevc_residual_bins<WRITE>(bs, lv + z * 64, 3, 0);
        } else if (part == 2) {
            if ((M.cbf_c >> z) & 1) hevc_residual_bins<WRITE>(bs, lv + 256 + z * 16, 2, 1);
        } else {
            if ((M.cbf_c >> (4 + z)) & 1) hevc_residual_bins<WRITE>(bs, lv + 320 + z * 16, 2, 2);
        }
    } else if (lane == 17) {
        bs.term(M.last_in_slice);
    }
}

__global__ void __launch_bounds__(CB_WARPS * 32) hevc_bins_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ __align__(16) int16_t lvs[CB_WARPS][384];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * CB_WARPS + warp;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    const int n = vcp_frame_of(s, gi);
    const size_t o = (size_t)gi * g.nmb + mbi;
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    const int sl = vcp_row_slice(b, my), row0 = vcp_row_first(b, my);
    CuCtx M;
    M.type = b.mbtype[o];
    { const int c = b.cbp[o]; M.cbf_y = c & 15; M.merge = (c >> 4) & 1; M.mvp_idx = (c >> 5) & 1; }
    M.cbf_c = b.modes[o]; M.mvd = b.mvd[o];
    M.tA = mx > 0 ? b.mbtype[o - 1] : -1;
    M.tB = my > row0 ? b.mbtype[o - g.mbw] : -1;
    M.idr = s.t == 0;
    {
        const int r1 = sl + 1 < g.slices ? vcp_slice_first_row(sl + 1, g.slices, g.mbh) : g.mbh;
        M.last_in_slice = (my == r1 - 1) && (mx == g.mbw - 1);
    }
    M.sao = -1; M.has_left = mx > 0; M.has_up = my > row0;
    M.sao_off[0] = M.sao_off[1] = M.sao_off[2] = M.sao_off[3] = 0;
    if (g.hevc_sao) {
        const uint8_t* rec = b.nnz + o * 24;
        M.sao = rec[0];
        for (int k = 0; k < 4; k++) M.sao_off[k] = (int)(int8_t)rec[1 + k];
    }
    if (M.type != 2)
        for (int i = lane; i < 48; i += 32) reinterpret_cast<uint4*>(lvs[warp])[i] = reinterpret_cast<const uint4*>(b.levels + o * VCP_LV_STRIDE)[i];
    __syncwarp();
    BinSink<false> cnt{nullptr, 0};
    hevc_cu_bins<false>(cnt, lvs[warp], M, lane);
    int incl = cnt.n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long off = 0;
    if (lane == 0) {
        off = atomicAdd(b.bins_cursor, (unsigned long long)total);
        if (off + (unsigned long long)total > b.bins_cap) { atomicExch(b.error_flag, 3); off = ~0ull; }
        else {
            b.mbdesc[(size_t)n * g.nmb + mbi] = make_uint2((uint32_t)off, (uint32_t)total | ((uint32_t)(off >> 32) << 20));
            atomicAdd(&b.slice_bins[(size_t)n * g.slices + sl], (uint32_t)total);
        }
    }
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off == ~0ull) return;
    BinSink<true> wr{b.bins + off + (incl - cnt.n), 0};
    hevc_cu_bins<true>(wr, lvs[warp], M, lane);
}

// slice_segment_header (7.3.6.1) for the stream structure of k6_hevc.cu, byte_alignment() included
__device__ __forceinline__ void hevc_slice_header(const VcpGeom& g, int first_ctb, bool idr, int poc, int qp, SeqBits& w) {
    w.put(1, first_ctb == 0);
    if (idr) w.put(1, 0);                  // no_output_of_prior_pics_flag
    w.ue(0);                               // slice_pic_parameter_set_id
    if (first_ctb) {
        int bits = 0;
        while ((1 << bits) < g.nmb) bits++;
        w.put(bits, (uint32_t)first_ctb);  // slice_segment_address
    }
    w.ue(idr ? 2 : 1);                     // slice_type
    if (!idr) {
        w.put(8, (uint32_t)(poc & 255));   // slice_pic_order_cnt_lsb
        w.put(1, 1);                       // short_term_ref_pic_set_sps_flag
    }
    if (g.hevc_sao) { w.put(1, 1); w.put(1, 0); }   // slice_sao_luma_flag, slice_sao_chroma_flag
    if (!idr) {
        w.put(1, 0);                       // num_ref_idx_active_override_flag
        w.ue(4);                           // five_minus_max_num_merge_cand
    }
    w.se(qp - 26);
    w.put(1, 1);
    while (w.pos & 7) w.pos++;             // the buffer is zero-filled
}

// ---- slice streams -------------------------------------------------------------------------------
// cabac_bins_kernel bump-allocates a macroblock's bins wherever the arena cursor stands, so the bins of a slice
// are scattered.  The arithmetic coder wants ONE contiguous, 16-byte aligned stream per slice (its lanes read
// their streams with 128-bit loads, a window ahead): this kernel lays the macroblocks' bins out in slice order.
// One CTA per (picture, slice) of the step: block scan of the macroblock counts, then warp-per-macroblock copies.
constexpr int GA_THREADS = 256;

__global__ void __launch_bounds__(GA_THREADS) cabac_gather_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ uint32_t wsum[GA_THREADS / 32];
    __shared__ uint32_t offs[GA_THREADS];
    __shared__ uint2 descs[GA_THREADS];
    __shared__ unsigned long long base_sh;
    const int S = g.slices;
    const int sl = blockIdx.x % S, gi = s.g0 + blockIdx.x / S;
    const int n = vcp_frame_of(s, gi);
    if (n >= s.nframes) return;
    const int r0 = vcp_slice_first_row(sl, S, g.mbh);
    const int r1 = sl + 1 < S ? vcp_slice_first_row(sl + 1, S, g.mbh) : g.mbh;
    const int first = r0 * g.mbw, count = (r1 - r0) * g.mbw;
    const uint2* desc = b.mbdesc + (size_t)n * g.nmb + first;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0 && *b.error_flag) base_sh = ~0ull;   // an arena overflowed earlier in this pass: descriptors are not trusted, the pass is run again
    else if (threadIdx.x == 0) {
        const uint32_t nb = b.slice_bins[(size_t)n * S + sl];
        // room for the stream, rounded up to whole 16-byte windows, plus one window the coder may read ahead
        const unsigned long long need = (((unsigned long long)nb + 7) & ~7ull) + 8;
        unsigned long long o = atomicAdd(b.sbins_cursor, need);
        if (o + need > b.sbins_cap) { atomicExch(b.error_flag, 3); o = ~0ull; }
        b.sslice_off[(size_t)n * S + sl] = o;
        base_sh = o;
    }
    __syncthreads();
    if (base_sh == ~0ull) { if (threadIdx.x == 0) b.sslice_off[(size_t)n * S + sl] = ~0ull; return; }
    uint16_t* dst = b.sbins + base_sh;
    uint32_t run = 0;
    for (int m0 = 0; m0 < count; m0 += GA_THREADS) {
        const int m = m0 + (int)threadIdx.x;
        const uint2 d = m < count ? desc[m] : make_uint2(0, 0);
        const uint32_t cnt = d.y & 0xfffff;
        uint32_t incl = cnt;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, k); if (lane >= k) incl += v; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t pre = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < GA_THREADS / 32; k++) { const uint32_t v = wsum[k]; if (k < warp) pre += v; tot += v; }
        offs[threadIdx.x] = run + pre + incl - cnt;
        descs[threadIdx.x] = d;
        __syncthreads();
        const int mend = min(GA_THREADS, count - m0);
        for (int j = warp; j < mend; j += GA_THREADS / 32) {
            const uint2 dj = descs[j];
            const uint32_t c = dj.y & 0xfffff;
            const uint16_t* src = b.bins + ((((unsigned long long)(dj.y >> 20)) << 32) | dj.x);
            uint16_t* q = dst + offs[j];
            for (uint32_t e = lane; e < c; e += 32) q[e] = src[e];
        }
        run += tot;
        __syncthreads();
    }
}

// ---- arithmetic coder: one LANE per slice ------------------------------------------------------------
// The coder of a slice is a serial dependency chain (state -> rLPS -> range/low -> renormalisation).  A warp that
// runs one chain pays an issue slot per scalar instruction (~90 per bin, measured round 1: the coder took a quarter of
// the step's issue capacity).  Here the 32 lanes of a warp run 32 chains in lock-step: the same instruction stream
// codes 32 bins, decision / bypass / terminate by predication, so a bin costs ~3 issue slots.
//   * lanes of a warp = slices of the SAME position in the GOP (all GOPs of the group x all slices of that picture):
//     their bin counts are alike, so the lanes end together (an IDR slice has ten times the bins of a P slice);
//   * every lane reads ITS contiguous stream (cabac_gather_kernel) with 128-bit loads, one window ahead;
//   * context states: one byte per context per lane in shared memory, [ctx][lane]; the rLPS / next-state tables are
//     replicated per lane ([state][lane]) so that a lookup never conflicts;
//   * bytes leave through per-lane byte stores into the slice's own output region (cap = one byte per bin + header:
//     a bin renormalises by at most 6 bits, so the region cannot overflow).
constexpr int AC_WARPS = 2;          // warps (32 slices each) per CTA
constexpr int AC_NCTX = 460;         // >= HC_NCTX

struct __align__(16) AcShared {
    uint32_t rtab[64][32];           // four rLPS bytes of a state, one copy per lane
    uint8_t ntab[64][32];            // next state on LPS
    uint8_t ctx[AC_WARPS][AC_NCTX][32];   // pStateIdx << 1 | valMPS
};

// Per-lane coder state.  `low` is kept 64 bits wide so that the output leaves 32 bits at a time (a flush every ~50
// bins instead of every ~13 with bytes: the flush block is the only divergent part of the loop).  Same arithmetic as
// the byte-wise low / queue / outstanding form of the oracle (cabac_encode / cabac_putbyte): the number written is the
// same, so the bytes are.  queue = bits shifted in since the last flush - 33 (the first output bit is dropped: 9.3.4.2's
// firstBitFlag); a flush takes 32 bits plus the carry above them.
struct LaneCoder {
    unsigned long long low;
    uint32_t range;
    int queue;
    uint32_t pend; bool have_pend;   // last word, not stored yet: a later carry may still increment it
    uint32_t outw;                   // outstanding 0xffffffff words behind it
    uint8_t* dst; uint32_t wpos;
    __device__ __forceinline__ void store4(uint32_t w) {   // big-endian, any alignment
        dst[wpos] = (uint8_t)(w >> 24); dst[wpos + 1] = (uint8_t)(w >> 16); dst[wpos + 2] = (uint8_t)(w >> 8); dst[wpos + 3] = (uint8_t)w;
        wpos += 4;
    }
    __device__ __forceinline__ void resolve(uint32_t carry) {   // the words held back become final
        if (have_pend) store4(pend + carry);
        for (; outw > 0; outw--) store4(carry ? 0u : 0xffffffffu);
        have_pend = false;
    }
    __device__ __forceinline__ void flush32() {
        const unsigned long long o = low >> (queue + 10);      // 32 bits + carry
        low &= (0x400ull << queue) - 1;
        queue -= 32;
        const uint32_t w = (uint32_t)o;
        if (w == 0xffffffffu) { outw++; return; }
        resolve((uint32_t)(o >> 32));
        pend = w; have_pend = true;
    }
};

// batch = pictures at GOP positions [t0, t1) of GOPs [g0, g0 + ngop); lane id = ((t - t0) * ngop + gop) * S + slice
__global__ void __launch_bounds__(AC_WARPS * 32) cabac_encode_kernel(VcpGeom g, VcpBufs b, VcpStep s, int t0, int t1) {
    extern __shared__ __align__(16) uint8_t ac_raw[];
    AcShared& A = *reinterpret_cast<AcShared*>(ac_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
        const int st = i >> 5;
        A.rtab[st][i & 31] = (uint32_t)vcp_cabac_range_lps[st][0] | ((uint32_t)vcp_cabac_range_lps[st][1] << 8) |
                             ((uint32_t)vcp_cabac_range_lps[st][2] << 16) | ((uint32_t)vcp_cabac_range_lps[st][3] << 24);
        A.ntab[st][i & 31] = (uint8_t)vcp_cabac_trans_lps[st];
    }
    __syncthreads();
    const int S = g.slices, nt = t1 - t0;
    const int id = (blockIdx.x * AC_WARPS + warp) * 32 + lane;
    const int total = s.ngop * nt * S;
    const int sl = id % S, gl = (id / S) % s.ngop, t = t0 + id / (S * s.ngop), gi = s.g0 + gl;
    const int n = gi * s.gop + t;
    bool live = id < total && n < s.nframes && *b.error_flag == 0;
    const size_t si = live ? (size_t)n * S + sl : 0;
    const bool idr = t == 0;
    const int qp = live ? b.qp[n] : 26;
    uint8_t (*ctx)[32] = A.ctx[warp];
    {   // context initialisation (9.3.1.1): every lane its own column
        const int tab = idr ? 0 : 1;
        const int nctx = g.hevc ? (int)HC_NCTX : NCTX;
        for (int i = 0; i < nctx; i++) {
            int m, nn;
            if (g.hevc) {   // H.265 9.3.2.2: slope / offset nibbles of initValue
                const int v = hevc_init_values[tab][i];
                m = (v >> 4) * 5 - 45; nn = ((v & 15) << 3) - 16;
            } else { m = vcp_cabac_init_mn[tab][i][0]; nn = vcp_cabac_init_mn[tab][i][1]; }
            const int pre = vcp_clip3(1, 126, ((m * vcp_clip3(0, 51, qp)) >> 4) + nn);
            ctx[i][lane] = (uint8_t)(pre <= 63 ? ((63 - pre) << 1) : (((pre - 64) << 1) | 1));
        }
    }
    const uint32_t nb = live ? b.slice_bins[si] : 0u;
    const unsigned long long soff = live ? b.sslice_off[si] : 0ull;
    if (soff == ~0ull) live = false;
    // output region: one byte per bin bounds what the coder can produce; + slice header and flush
    const unsigned long long cap = ((unsigned long long)nb + 128 + 15) & ~15ull;
    LaneCoder C;
    C.low = 0; C.range = 510; C.queue = -33; C.pend = 0; C.have_pend = false; C.outw = 0; C.dst = nullptr; C.wpos = 0;
    unsigned long long obase = 0;
    if (live) {
        obase = atomicAdd(b.crbsp_cursor, cap);
        if (obase + cap > b.crbsp_cap) { atomicExch(b.error_flag, 4); b.cslice_bytes[si] = 0; live = false; }
    }
    if (live) {
        const int r0 = vcp_slice_first_row(sl, S, g.mbh);
        uint8_t* dst = b.crbsp + obase;
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(dst)[1] = make_uint4(0, 0, 0, 0);
        SeqBits w{dst, 0};
        if (g.hevc) hevc_slice_header(g, r0 * g.mbw, idr, t, qp, w);
        else {
            slice_header_bits(g, r0 * g.mbw, idr, t, (s.gop0 + gi) & 1, qp, &w);
            while (w.pos & 7) w.put(1, 1);   // cabac_alignment_one_bit
        }
        C.dst = dst; C.wpos = w.pos >> 3;
    }
    const uint4* win = reinterpret_cast<const uint4*>(b.sbins + (live ? soff : 0ull));
    const uint32_t mybins = live ? nb : 0u;
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (mybins) cur = __ldg(win);
    bool ended = false;               // the end_of_slice terminate bin (value 1) was met: always the last bin of a stream
    // Software pipeline: while bin j runs through the range / low chain, the context state of bin j+1 is already being
    // fetched (before bin j's state is written back: if both use the same context the fetched byte is stale and the
    // register value replaces it) and its table entries follow as soon as that state is known.  The state machine does
    // not depend on range or low, so only the ~10 dependent instructions of the range update separate two bins.
    uint32_t bin = cur.x & 0xffffu;
    uint32_t st = ctx[bin & 1023u][lane];
    uint32_t rl4 = A.rtab[st >> 1][lane], nlps = A.ntab[st >> 1][lane];
    for (uint32_t k0 = 0; __any_sync(0xffffffffu, k0 < mybins); k0 += 8) {
        // the next window is in flight while this one is coded (streams are padded by one window)
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (k0 + 8 < mybins) nxt = __ldg(win + (k0 >> 3) + 1);
        const uint32_t w4[5] = {cur.x, cur.y, cur.z, cur.w, nxt.x};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            // One bin per lane, decision / bypass / terminate(0) by selection, no branches.  A lane past its stream (or at
            // its final terminate bin) runs the same instructions with "nothing": range stays, nothing is added or shifted.
            const uint32_t bin_n = (w4[(j + 1) >> 1] >> (16 * ((j + 1) & 1))) & 0xffffu;
            const uint32_t c = bin & 1023u, c_n = bin_n & 1023u;       // 0 for bypass / terminate bins: a harmless read
            const uint32_t ld_n = ctx[c_n][lane];
            const bool act = k0 + j < mybins;
            const uint32_t val = (bin >> 10) & 1u;
            const bool is_byp = act && (bin & BIN_BYPASS);
            const bool is_term = act && (bin & BIN_TERM);
            const bool is_dec = act && !(bin & (BIN_BYPASS | BIN_TERM));
            if (is_term && val) ended = true;
            const uint32_t ps = st >> 1, mps = st & 1u;
            const bool lps = val != mps;
            const uint32_t nst = lps ? ((nlps << 1) | (ps == 0 ? mps ^ 1u : mps)) : (((ps < 62 ? ps + 1 : ps) << 1) | mps);
            if (is_dec) ctx[c][lane] = (uint8_t)nst;
            const uint32_t st_n = (is_dec && c_n == c) ? nst : ld_n;
            const uint32_t rl4_n = A.rtab[st_n >> 1][lane], nlps_n = A.ntab[st_n >> 1][lane];
            // range / low: decision: LPS takes rLPS and adds the MPS range to low; bypass: low doubles first, then adds
            // range for a 1; terminate(0): range loses 2
            const uint32_t range = C.range;
            const uint32_t rlps = __byte_perm(rl4, 0u, 0x4440u | ((range >> 6) & 3u));
            const uint32_t rmps = range - rlps;
            uint32_t r1 = range, add = 0u;
            if (is_dec) { r1 = lps ? rlps : rmps; add = lps ? rmps : 0u; }
            if (is_byp) add = val ? range : 0u;
            if (is_term && !val) r1 = range - 2u;
            const int sh = __clz(r1) - 23;                       // 0 whenever the range did not shrink below 256
            const int tot = sh + (is_byp ? 1 : 0);
            C.low = (C.low << tot) + ((unsigned long long)(add << sh));
            C.range = r1 << sh;
            C.queue += tot;
            if (C.queue >= 0) C.flush32();
            bin = bin_n; st = st_n; rl4 = rl4_n; nlps = nlps_n;
        }
        cur = nxt;
    }
    if (live) {
        if (!ended) { atomicExch(b.error_flag, 5); b.cslice_bytes[si] = 0; return; }   // a stream that does not end in end_of_slice: never produced
        // end of slice: flush (9.3.4.5), stop bit included -- range was not touched by the final bin above
        C.range -= 2;
        C.low += C.range;
        C.low = (C.low << 10) | 0x400ull;
        C.queue += 10;
        if (C.queue >= 0) C.flush32();
        // what is left above the register's ready boundary: r = queue + 32 bits (0..31) and the carry over them
        const int r = C.queue + 32;
        const unsigned long long o = C.low >> 10;
        C.resolve((uint32_t)(o >> r) & 1u);
        int left = r;
        for (; left >= 8; left -= 8) C.dst[C.wpos++] = (uint8_t)(o >> (left - 8));
        if (left > 0) C.dst[C.wpos++] = (uint8_t)((o & ((1ull << left) - 1)) << (8 - left));
        if ((unsigned long long)C.wpos > cap) { atomicExch(b.error_flag, 4); b.cslice_bytes[si] = 0; }   // cannot happen: cap bounds the coder's output
        else { b.cslice_bytes[si] = C.wpos; b.cslice_off[si] = obase; }
    }
}

// grid: x = (GOP, t, slice) of the batch
__global__ void __launch_bounds__(PACK_THREADS) cabac_pack_kernel(VcpGeom g, VcpBufs b, VcpStep s, int t0, int t1) {
    const int S = g.slices, nt = t1 - t0;
    const int id = blockIdx.x;
    const int sl = id % S, t = t0 + (id / S) % nt, gi = s.g0 + id / (S * nt);
    const int n = gi * s.gop + t;
    if (n >= s.nframes) return;
    const uint32_t bytes = b.cslice_bytes[(size_t)n * S + sl];
    if (!bytes) return;
    nal_pack_body(g, b, b.crbsp + b.cslice_off[(size_t)n * S + sl], bytes, n, sl, t == 0);
}

// frame_bits estimate for rate control: bins * VCP_CABAC_BITS_PER_BIN_Q4 / 16, per slice like the oracle
__global__ void cabac_rc_bits_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= s.ngop) return;
    const int n = vcp_frame_of(s, s.g0 + k);
    uint32_t bits = 0;
    for (int sl = 0; sl < g.slices; sl++)
        bits += (uint32_t)(((unsigned long long)b.slice_bins[(size_t)n * g.slices + sl] * VCP_CABAC_BITS_PER_BIN_Q4) >> 4);
    b.frame_bits[n] = bits;
}

}  // namespace

void vcp_launch_cabac_bins(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + CB_WARPS - 1) / CB_WARPS, s.ngop);
    cabac_bins_kernel<<<grid, CB_WARPS * 32, 0, st>>>(g, b, s);
    static const int dbg_skip = [] { const char* e = getenv("VCPENC_DEBUG_SKIP_CODER"); return e ? atoi(e) : 0; }();   // timing experiments only
    if (dbg_skip < 2) cabac_gather_kernel<<<s.ngop * g.slices, GA_THREADS, 0, st>>>(g, b, s);
    if (g.rc_fb) cabac_rc_bits_kernel<<<(s.ngop + 63) / 64, 64, 0, st>>>(g, b, s);
}

void vcp_launch_hevc_bins(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + CB_WARPS - 1) / CB_WARPS, s.ngop);
    hevc_bins_kernel<<<grid, CB_WARPS * 32, 0, st>>>(g, b, s);
    cabac_gather_kernel<<<s.ngop * g.slices, GA_THREADS, 0, st>>>(g, b, s);
    if (g.rc_fb) cabac_rc_bits_kernel<<<(s.ngop + 63) / 64, 64, 0, st>>>(g, b, s);
}

void vcp_launch_cabac_encode(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, int t0, int t1, cudaStream_t st) {
    const int total = s.ngop * (t1 - t0) * g.slices;
    if (total <= 0) return;
    // function attributes are per device: a process may drive several GPUs from different threads
    cudaFuncSetAttribute(cabac_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AcShared));
    cabac_encode_kernel<<<(total + AC_WARPS * 32 - 1) / (AC_WARPS * 32), AC_WARPS * 32, sizeof(AcShared), st>>>(g, b, s, t0, t1);
    cabac_pack_kernel<<<total, PACK_THREADS, 0, st>>>(g, b, s, t0, t1);
}
