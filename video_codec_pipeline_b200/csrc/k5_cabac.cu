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

// ---- arithmetic coder, one WARP per slice ------------------------------------------------------
// The coder is a serial dependency chain (state -> rLPS -> range/low -> renormalisation), so what
// bounds a slice is latency per bin, not throughput.  Lane 0 runs the chain out of shared memory;
// the other 31 lanes exist to keep it fed: they gather the macroblocks' bins into a shared ring
// with coalesced loads and drain the produced bytes with coalesced stores.  Many such warps share
// an SM, so the machine interleaves as many independent chains as there are slices in the batch.
constexpr int AC_WARPS = 4;          // slices per CTA
constexpr int BINBUF = 2048;         // bins staged per round
constexpr int OUTBUF = BINBUF + 32;  // a bin renormalises by at most 7 bits

struct ArithCoder {
    uint32_t low, range;
    int queue, outstanding, last;   // last: pending byte not yet stored (-1: none)
    uint8_t* out;                   // shared staging of this round
    int nout;
    bool ovf;                       // a run of outstanding 0xff bytes longer than the staging (never seen; reported)
    __device__ __forceinline__ void store(int v) { if (nout < OUTBUF) out[nout++] = (uint8_t)v; else ovf = true; }
    __device__ __forceinline__ void emit(int o) {   // 8 bits + carry in bit 8
        if ((o & 0xff) == 0xff) { outstanding++; return; }
        const int carry = o >> 8;
        if (last >= 0) store(last + carry);
        while (outstanding > 0) { store(carry ? 0x00 : 0xff); outstanding--; }
        last = o & 0xff;
    }
    __device__ __forceinline__ void putbyte() {
        if (queue >= 0) {
            const int o = (int)(low >> (queue + 10));
            low &= (0x400u << queue) - 1;
            queue -= 8;
            emit(o);
        }
    }
};

// A context's record is a copy of the table entry of its probability state, so the range/low
// chain never waits for a table lookup: x = the four rLPS bytes, y = next state on MPS | next state
// on LPS << 8 | valMPS << 16.  The lookup of the successor entry runs beside the chain.
struct __align__(16) AcScratch {
    uint2 rec[NCTX];
    uint16_t bins[BINBUF + 2];
    uint8_t out[OUTBUF];
};

// batch = pictures at GOP positions [t0, t1) of GOPs [g0, g0 + ngop); one warp per (GOP, t, slice)
__global__ void __launch_bounds__(AC_WARPS * 32) cabac_encode_kernel(VcpGeom g, VcpBufs b, VcpStep s, int t0, int t1) {
    __shared__ AcScratch scr[AC_WARPS];
    __shared__ uint2 ent[128];      // per state<<1|mps: x = four rLPS bytes, y = next on MPS | next on LPS << 8 | valMPS << 16
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        const int st = i >> 1, mps = i & 1;
        const uint32_t l4 = (uint32_t)vcp_cabac_range_lps[st][0] | ((uint32_t)vcp_cabac_range_lps[st][1] << 8) |
                            ((uint32_t)vcp_cabac_range_lps[st][2] << 16) | ((uint32_t)vcp_cabac_range_lps[st][3] << 24);
        const uint32_t nm = (uint32_t)(((st < 62 ? st + 1 : st) << 1) | mps);
        const uint32_t nl = (uint32_t)((vcp_cabac_trans_lps[st] << 1) | (st == 0 ? mps ^ 1 : mps));
        ent[i] = make_uint2(l4, nm | (nl << 8) | ((uint32_t)mps << 16));
    }
    __syncthreads();
    const int S = g.slices, nt = t1 - t0;
    const int id = blockIdx.x * AC_WARPS + warp;
    if (id >= s.ngop * nt * S) return;
    const int sl = id % S, t = t0 + (id / S) % nt, gi = s.g0 + id / (S * nt);
    const int n = gi * s.gop + t;
    if (n >= s.nframes) return;
    const bool idr = t == 0;
    const int qp = b.qp[n];
    AcScratch& A = scr[warp];
    {   // context initialisation (9.3.1.1), lanes share the contexts
        const int tab = idr ? 0 : 1;
        for (int i = lane; i < (g.hevc ? (int)HC_NCTX : NCTX); i += 32) {
            int m, nn;
            if (g.hevc) {   // H.265 9.3.2.2: slope / offset nibbles of initValue
                const int v = hevc_init_values[tab][i];
                m = (v >> 4) * 5 - 45; nn = ((v & 15) << 3) - 16;
            } else { m = vcp_cabac_init_mn[tab][i][0]; nn = vcp_cabac_init_mn[tab][i][1]; }
            const int pre = vcp_clip3(1, 126, ((m * vcp_clip3(0, 51, qp)) >> 4) + nn);
            A.rec[i] = ent[pre <= 63 ? ((63 - pre) << 1) : (((pre - 64) << 1) | 1)];
        }
    }
    const int r0 = vcp_slice_first_row(sl, S, g.mbh);
    const int r1 = sl + 1 < S ? vcp_slice_first_row(sl + 1, S, g.mbh) : g.mbh;
    const int first = r0 * g.mbw, count = (r1 - r0) * g.mbw;
    // output region: slice header + 4 bits per bin is more than the coder can produce on average
    const uint32_t nb = b.slice_bins[(size_t)n * S + sl];
    const unsigned long long cap = ((unsigned long long)nb / 2 + 96 + 15) & ~15ull;
    unsigned long long base = 0;
    if (lane == 0) {
        base = atomicAdd(b.crbsp_cursor, cap);
        if (base + cap > b.crbsp_cap) { atomicExch(b.error_flag, 4); b.cslice_bytes[(size_t)n * S + sl] = 0; base = ~0ull; }
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base == ~0ull) return;
    uint8_t* dst = b.crbsp + base;
    uint32_t wpos = 0;   // bytes written to dst (lane 0's view is broadcast)
    if (lane == 0) {
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(dst)[1] = make_uint4(0, 0, 0, 0);
        SeqBits w{dst, 0};
        if (g.hevc) hevc_slice_header(g, first, idr, t, qp, w);
        else {
            slice_header_bits(g, first, idr, t, (s.gop0 + gi) & 1, qp, &w);
            while (w.pos & 7) w.put(1, 1);   // cabac_alignment_one_bit
        }
        wpos = w.pos >> 3;
    }
    wpos = __shfl_sync(0xffffffffu, wpos, 0);
    __syncwarp();
    ArithCoder C;
    C.low = 0; C.range = 510; C.queue = -9; C.outstanding = 0; C.last = -1; C.out = A.out; C.nout = 0; C.ovf = false;
    const uint2* desc = b.mbdesc + (size_t)n * g.nmb + first;
    int mb = 0;           // next macroblock to stage
    uint32_t part = 0;    // bins of macroblock `mb` already consumed (macroblocks larger than the ring)
    bool overflow = false;
    while (mb < count) {
        // ---- stage: lanes look at the next 32 macroblocks, take as many as fit --------------------
        uint2 d = make_uint2(0, 0);
        if (mb + lane < count) d = desc[mb + lane];
        uint32_t cnt = d.y & 0xfffff;
        if (lane == 0) cnt -= part;
        uint32_t incl = cnt;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, k);
            if (lane >= k) incl += v;
        }
        const uint32_t fits = __ballot_sync(0xffffffffu, incl <= (uint32_t)BINBUF && mb + lane < count);
        int take = __ffs(~fits) - 1;                 // leading macroblocks that fit entirely
        if (take < 0) take = 32;
        uint32_t nstaged;
        if (take == 0) {                             // one macroblock larger than the ring: take a piece
            const unsigned long long off = (((unsigned long long)(__shfl_sync(0xffffffffu, d.y, 0) >> 20)) << 32) | __shfl_sync(0xffffffffu, d.x, 0);
            for (int e = lane; e < BINBUF; e += 32) A.bins[e] = b.bins[off + part + e];
            nstaged = BINBUF;
            part += BINBUF;
        } else {
            for (int j = 0; j < take; j++) {
                const uint32_t dx = __shfl_sync(0xffffffffu, d.x, j), dy = __shfl_sync(0xffffffffu, d.y, j);
                const uint32_t c = __shfl_sync(0xffffffffu, cnt, j), o = __shfl_sync(0xffffffffu, incl, j) - c;
                const uint16_t* src = b.bins + ((((unsigned long long)(dy >> 20)) << 32) | dx) + (j == 0 ? part : 0u);
                for (uint32_t e = lane; e < c; e += 32) A.bins[o + e] = src[e];
            }
            nstaged = __shfl_sync(0xffffffffu, incl, take - 1);
            mb += take;
            part = 0;
        }
        __syncwarp();
        // ---- the chain: lane 0 ----------------------------------------------------------------------
        if (lane == 0) {
            C.nout = 0;
            // software pipeline: the next bin and its context record are fetched while the current
            // bin is coded; a repeated context takes the freshly updated record instead.  Bypass and
            // terminate bins carry ctx 0, so the record fetch needs no test.
            A.bins[nstaged] = 0;
            uint32_t cur = A.bins[0];
            uint2 rc = A.rec[cur & 1023u];
            uint32_t low = C.low, range = C.range;
            int queue = C.queue;
#pragma unroll 2
            for (uint32_t k = 0; k < nstaged; k++) {
                const uint32_t nxt = A.bins[k + 1];
                uint2 rn = A.rec[nxt & 1023u];
                if (cur & (BIN_BYPASS | BIN_TERM)) {
                    if (cur & BIN_BYPASS) {
                        low = (low << 1) + ((cur & 0x400u) ? range : 0u);
                        queue += 1;
                    } else {
                        range -= 2;
                        if (cur & 0x400u) {   // end of slice: flush (9.3.4.5), stop bit included
                            low += range;
                            C.low = low << 7; C.range = 2u << 7; C.queue = queue + 7;
                            C.putbyte();
                            low = (C.low << 3) | 0x400u; queue = C.queue + 3; range = C.range;
                        } else {
                            const int sh = __clz(range) - 23;
                            range <<= sh; low <<= sh; queue += sh;
                        }
                    }
                } else {
                    const uint32_t rlps = __byte_perm(rc.x, 0u, 0x4440u | ((range >> 6) & 3u));
                    const uint32_t lps = ((cur >> 10) ^ (rc.y >> 16)) & 1u;
                    const uint2 nrec = ent[__byte_perm(rc.y, 0u, 0x4440u + lps)];   // successor entry, beside the chain
                    range -= rlps;
                    if (lps) { low += range; range = rlps; }
                    A.rec[cur & 1023u] = nrec;
                    if (((nxt ^ cur) & 1023u) == 0u) rn = nrec;    // (a special next bin has ctx 0 != ctx of a decision... unless ctx 0: never coded)
                    const int sh = __clz(range) - 23;
                    range <<= sh; low <<= sh; queue += sh;
                }
                if (queue >= 0) {
                    C.low = low; C.queue = queue;
                    C.putbyte();
                    low = C.low; queue = C.queue;
                }
                cur = nxt; rc = rn;
            }
            C.low = low; C.range = range; C.queue = queue;
            if (mb >= count && part == 0) {
                // remaining bits above the register's ready boundary, then the pending bytes
                const int r = C.queue + 8;                      // 0..7 bits left
                const int o = (int)(C.low >> 10);
                const int carry = o >> r;
                if (C.last >= 0) C.store(C.last + carry);
                while (C.outstanding > 0) { C.store(carry ? 0x00 : 0xff); C.outstanding--; }
                if (r > 0) C.store((o & ((1 << r) - 1)) << (8 - r));
            }
        }
        __syncwarp();
        // ---- drain: all lanes ---------------------------------------------------------------------
        const int nout = __shfl_sync(0xffffffffu, C.nout, 0);
        if (__shfl_sync(0xffffffffu, (int)C.ovf, 0) || (unsigned long long)wpos + (unsigned)nout > cap) overflow = true;
        else for (int e = lane; e < nout; e += 32) dst[wpos + e] = A.out[e];
        wpos += (uint32_t)nout;
        __syncwarp();
    }
    if (lane == 0) {
        if (overflow) { atomicExch(b.error_flag, 4); b.cslice_bytes[(size_t)n * S + sl] = 0; }
        else { b.cslice_bytes[(size_t)n * S + sl] = wpos; b.cslice_off[(size_t)n * S + sl] = base; }
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
    if (g.rc_abr) cabac_rc_bits_kernel<<<(s.ngop + 63) / 64, 64, 0, st>>>(g, b, s);
}

void vcp_launch_hevc_bins(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + CB_WARPS - 1) / CB_WARPS, s.ngop);
    hevc_bins_kernel<<<grid, CB_WARPS * 32, 0, st>>>(g, b, s);
    if (g.rc_abr) cabac_rc_bits_kernel<<<(s.ngop + 63) / 64, 64, 0, st>>>(g, b, s);
}

void vcp_launch_cabac_encode(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, int t0, int t1, cudaStream_t st) {
    const int total = s.ngop * (t1 - t0) * g.slices;
    if (total <= 0) return;
    cabac_encode_kernel<<<(total + AC_WARPS - 1) / AC_WARPS, AC_WARPS * 32, 0, st>>>(g, b, s, t0, t1);
    cabac_pack_kernel<<<total, PACK_THREADS, 0, st>>>(g, b, s, t0, t1);
}
