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
//   cabac_encode_kernel : for a batch of pictures of every resident GOP, one LANE per slice walks
//                         its macroblocks' bins through the arithmetic coder (byte-wise
//                         low/queue/outstanding form, bit-identical to 9.3.4.2's PutBit procedure)
//                         and writes the slice RBSP.  32 slices per warp; context states live in
//                         shared memory, column per lane.
//   cabac_pack_kernel   : NAL encapsulation of those RBSPs (shared with CAVLC, vcp_entropy.cuh).
//
// Replaces x264's cabac.c inside the ffmpeg child (/root/reference/cmd/consumer.go:376-382; the
// h264-cpu preset is High/CABAC by default).  Output bytes are identical to oracle/h264_oracle.c
// (write_slice_data_cabac, cabac_block, cabac_mvd, cabac_encode/bypass/terminate).
#include "vcp_dev.cuh"

#define VCP_TAB static __device__ const
#include "h264_cabac_tables.h"
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
    bool idr, last_in_slice;
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
        if (cbpl & (1 << (blk >> 2))) {
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

// ---- arithmetic coder, one lane per slice -------------------------------------------------------
struct ArithLane {
    uint32_t low, range;
    int queue, outstanding, last;   // last: pending byte not yet stored (-1: none)
    uint8_t* p;
    uint8_t* end;
    bool overflow;
    __device__ __forceinline__ void store(int v) { if (p < end) *p++ = (uint8_t)v; else overflow = true; }
    __device__ __forceinline__ void emit(int out) {   // out: 8 bits + carry in bit 8
        if ((out & 0xff) == 0xff) { outstanding++; return; }
        const int carry = out >> 8;
        if (last >= 0) store(last + carry);
        while (outstanding > 0) { store(carry ? 0x00 : 0xff); outstanding--; }
        last = out & 0xff;
    }
    __device__ __forceinline__ void putbyte() {
        if (queue >= 0) {
            const int out = (int)(low >> (queue + 10));
            low &= (0x400u << queue) - 1;
            queue -= 8;
            emit(out);
        }
    }
    __device__ __forceinline__ void renorm() {
        const int sh = __clz(range) - 23;     // range in [2, 510] -> bring bit 8 up
        range <<= sh; low <<= sh; queue += sh;
        putbyte();
    }
};

// batch = pictures at GOP positions [t0, t1) of GOPs [g0, g0 + ngop); one lane per (GOP, t, slice)
__global__ void __launch_bounds__(32) cabac_encode_kernel(VcpGeom g, VcpBufs b, VcpStep s, int t0, int t1) {
    __shared__ uint8_t state[NCTX][32];
    __shared__ uint32_t lps4[64];
    __shared__ uint8_t next_mps[128], next_lps[128];
    const int lane = threadIdx.x;
    for (int i = lane; i < 64; i += 32)
        lps4[i] = (uint32_t)vcp_cabac_range_lps[i][0] | ((uint32_t)vcp_cabac_range_lps[i][1] << 8) |
                  ((uint32_t)vcp_cabac_range_lps[i][2] << 16) | ((uint32_t)vcp_cabac_range_lps[i][3] << 24);
    for (int i = lane; i < 128; i += 32) {
        const int st = i >> 1, mps = i & 1;
        next_mps[i] = (uint8_t)(((st < 62 ? st + 1 : st) << 1) | mps);
        next_lps[i] = (uint8_t)((vcp_cabac_trans_lps[st] << 1) | (st == 0 ? mps ^ 1 : mps));
    }
    __syncwarp();
    const int S = g.slices, nt = t1 - t0;
    const int id = blockIdx.x * 32 + lane;
    if (id >= s.ngop * nt * S) return;
    const int sl = id % S, t = t0 + (id / S) % nt, gi = s.g0 + id / (S * nt);
    const int n = gi * s.gop + t;
    if (n >= s.nframes) return;
    const bool idr = t == 0;
    const int qp = b.qp[n];
    {   // context initialisation (9.3.1.1)
        const int tab = idr ? 0 : 1;
        for (int i = 0; i < NCTX; i++) {
            const int m = vcp_cabac_init_mn[tab][i][0], nn = vcp_cabac_init_mn[tab][i][1];
            const int pre = vcp_clip3(1, 126, ((m * vcp_clip3(0, 51, qp)) >> 4) + nn);
            state[i][lane] = pre <= 63 ? (uint8_t)((63 - pre) << 1) : (uint8_t)(((pre - 64) << 1) | 1);
        }
    }
    const int r0 = vcp_slice_first_row(sl, S, g.mbh);
    const int r1 = sl + 1 < S ? vcp_slice_first_row(sl + 1, S, g.mbh) : g.mbh;
    const int first = r0 * g.mbw, count = (r1 - r0) * g.mbw;
    // output region: slice header + 4 bits per bin is more than the coder can produce
    const uint32_t nb = b.slice_bins[(size_t)n * S + sl];
    const unsigned long long cap = ((unsigned long long)nb / 2 + 64 + 15) & ~15ull;
    const unsigned long long base = atomicAdd(b.crbsp_cursor, cap);
    if (base + cap > b.crbsp_cap) { atomicExch(b.error_flag, 4); b.cslice_bytes[(size_t)n * S + sl] = 0; return; }
    uint8_t* dst = b.crbsp + base;
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(0, 0, 0, 0);
    SeqBits w{dst, 0};
    slice_header_bits(g, first, idr, t, (s.gop0 + gi) & 1, qp, &w);
    while (w.pos & 7) w.put(1, 1);   // cabac_alignment_one_bit
    ArithLane A;
    A.low = 0; A.range = 510; A.queue = -9; A.outstanding = 0; A.last = -1;
    A.p = dst + (w.pos >> 3); A.end = dst + cap; A.overflow = false;
    const uint2* desc = b.mbdesc + (size_t)n * g.nmb + first;
    for (int i = 0; i < count; i++) {
        const uint2 d = desc[i];
        const uint16_t* bp = b.bins + (((unsigned long long)(d.y >> 20) << 32) | d.x);
        const int cnt = (int)(d.y & 0xfffff);
        for (int k = 0; k < cnt; k++) {
            const uint32_t v = bp[k];
            const int bin = (v >> 10) & 1;
            if (v & BIN_BYPASS) {
                A.low <<= 1;
                if (bin) A.low += A.range;
                A.queue += 1;
                A.putbyte();
            } else if (v & BIN_TERM) {
                A.range -= 2;
                if (bin) {   // end of slice: flush (9.3.4.5), stop bit included
                    A.low += A.range;
                    A.range = 2;
                    A.renorm();
                    A.low = (A.low << 3) | 0x400u;
                    A.queue += 3;
                    A.putbyte();
                } else A.renorm();
            } else {
                const int ctx = (int)(v & 1023);
                const int st = state[ctx][lane];
                const uint32_t rlps = (lps4[st >> 1] >> (((A.range >> 6) & 3) * 8)) & 255;
                A.range -= rlps;
                if (bin != (st & 1)) { A.low += A.range; A.range = rlps; state[ctx][lane] = next_lps[st]; }
                else state[ctx][lane] = next_mps[st];
                A.renorm();
            }
        }
    }
    // remaining bits above the register's ready boundary, then the pending bytes
    {
        const int r = A.queue + 8;                      // 0..7 bits left
        const int out = (int)(A.low >> 10);
        const int carry = out >> r;
        if (A.last >= 0) A.store(A.last + carry);
        while (A.outstanding > 0) { A.store(carry ? 0x00 : 0xff); A.outstanding--; }
        if (r > 0) A.store((out & ((1 << r) - 1)) << (8 - r));
    }
    if (A.overflow) { atomicExch(b.error_flag, 4); b.cslice_bytes[(size_t)n * S + sl] = 0; return; }
    b.cslice_bytes[(size_t)n * S + sl] = (uint32_t)(A.p - dst);
    b.cslice_off[(size_t)n * S + sl] = base;
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

void vcp_launch_cabac_encode(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, int t0, int t1, cudaStream_t st) {
    const int total = s.ngop * (t1 - t0) * g.slices;
    if (total <= 0) return;
    cabac_encode_kernel<<<(total + 31) / 32, 32, 0, st>>>(g, b, s, t0, t1);
    cabac_pack_kernel<<<total, PACK_THREADS, 0, st>>>(g, b, s, t0, t1);
}
