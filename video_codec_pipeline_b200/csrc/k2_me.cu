// K2 — integer motion estimation, search windows staged by TMA.
//
//  K2a  me_prepass_kernel : hierarchical full-pel search on the ORIGINAL frames.  It does not
//        depend on any reconstruction, so one launch covers every P-frame of every GOP that
//        is resident — this is where the bulk of the SAD work lives and it is off the
//        frame-to-frame dependency chain.
//          L1: half-res 8x8 block per macroblock, exhaustive +-12 (625 candidates), SAD via
//              VABSDIFF4.U8.ACC.  CTA = 16 macroblocks of a row.  One elected thread asks the copy
//              engine (cp.async.bulk.tensor) for the CTA's half-res window and its 16 current blocks;
//              the CTA then makes three more copies of the window, shifted by one, two and three bytes,
//              so that every candidate column is an ALIGNED word pair in the copy whose shift equals
//              its misalignment: the inner loop is two LDS.32 per reference row feeding up to 16
//              VABSDIFF4, no funnel shifts, no address arithmetic.  thread = (macroblock, dx) column,
//              laid out so that the two half-warps of a warp read disjoint banks and 400 of 416
//              launched lanes carry candidates.
//          L0: full-res 16x16, +-2 around twice the L1 vector: warp = macroblock, one TMA box for the
//              reference region and one for the current block; lane = (horizontal offset, 4-row group).
//  K2b  me_refine_kernel  : inside the per-frame chain, against the RECONSTRUCTED reference:
//        11 full-pel candidates, then 8 half-pel and 8 quarter-pel neighbours taken from the four
//        sample planes (G, b, h, j of 8.4.2.2.1, k2_hpel.cu).  One 4D TMA box (x, y, plane, slot)
//        brings the window of all four planes around the pre-pass vector; the zero and the predictor
//        candidates get their own G boxes in the same transaction.
//  TMA boxes start on 16-byte boundaries (vcp_tma.cuh): every window carries its misalignment 0..15.
//
// Replaces x264's `me=hex subme=7` / NVENC's ME inside the ffmpeg child
// (/root/reference/cmd/consumer.go:376-382).  Decisions are bit-identical to
// oracle/h264_oracle.c (me_prepass, me_refine_mb).
#include "vcp_dev.cuh"
#include "vcp_luma_interp.cuh"
#include "vcp_tma.cuh"
#include "vcp_hevc_qpel.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// K2a
// ---------------------------------------------------------------------------------------------
constexpr int PP_MBS = 16;                         // macroblocks per CTA (one row segment)
constexpr int PP_THREADS = 416;                    // 13 warps: 12.5 carry the L1 columns
constexpr int PP0_WARPS = 4;                       // L0 kernel: macroblocks (warps) per CTA
constexpr int PP_NDX = 2 * VCP_ME_R1 + 1;          // 25
constexpr int PP_ITEMS = PP_MBS * PP_NDX;          // 400 (macroblock, dx) columns
constexpr int PP_WIN_USED = 160;                   // bytes of a window row the columns read (the box brings 16 more for the shifts)

struct __align__(128) PrepassSmem {
    uint8_t win[4][VCP_L1_WIN_H][VCP_L1_WIN_W];    // copy k = the window shifted left by k bytes; copy 0 is where the box lands
    uint8_t hcur[8][VCP_L1_CUR_W];                 // the CTA's 16 half-res current blocks
    uint64_t bar_l1;
    uint32_t best1[PP_MBS];
};
static_assert(sizeof(PrepassSmem) <= 48 * 1024, "static shared memory, three CTAs per SM");
struct __align__(128) PrepassL0Warp {
    uint8_t ref[1024];                             // VCP_L0_REF_H rows of VCP_L0_REF_W bytes
    uint8_t cur[16][16];
    uint64_t bar;
};
static_assert(VCP_L0_REF_H * VCP_L0_REF_W <= 1024, "L0 reference box");

// grid z = frame (t < 0: every resident frame) or GOP (t >= 0: picture t of every GOP, so that the
// pre-pass of later pictures runs beside the reconstruction chain of earlier ones)
// Two kernels, one TMA phase each (a single kernel waited twice per CTA at 2 CTAs per SM: 45 % of its warp samples sat in
// the waits).  L1 leaves twice its best half-res vector in mvfp[], L0 replaces it by the full-res vector.
__global__ void __launch_bounds__(PP_THREADS, 3) me_prepass_kernel(VcpGeom g, VcpBufs b, int nframes, int gop, int t, int g0,
                                                                    const __grid_constant__ VcpTmaps tm) {
    __shared__ PrepassSmem S;
    const int tid = threadIdx.x;
    const int n = t < 0 ? (int)blockIdx.z : ((int)blockIdx.z + g0) * gop + t;
    if (n >= nframes || n % gop == 0) return;  // IDR: no search (uniform for the CTA)
    const int my = blockIdx.y, mx0 = blockIdx.x * PP_MBS;
    if (tid == 0) {
        mbar_init(&S.bar_l1, 1);
        mbar_init_fence();
    }
    if (tid < PP_MBS) S.best1[tid] = 0xffffffffu;
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&S.bar_l1, (uint32_t)(VCP_L1_WIN_H * VCP_L1_WIN_W + 8 * VCP_L1_CUR_W));
        // window: rows 8my-12 .. 8my+19, columns from 8mx0-16; plane coordinates carry the border (VCP_PAD1 = 16), and
        // mx0 is a multiple of 16, so both boxes start on a 128-byte boundary
        tma_load_3d(&S.win[0][0][0], &tm.h_win, &S.bar_l1, 8 * mx0 - 16 + VCP_PAD1, 8 * my - VCP_ME_R1 + VCP_PAD1, n - 1);
        tma_load_3d(&S.hcur[0][0], &tm.h_cur, &S.bar_l1, 8 * mx0 + VCP_PAD1, 8 * my + VCP_PAD1, n);
    }
    mbar_wait(&S.bar_l1, 0);
    // the three shifted copies: 16-byte units, five source words each
    for (int u = tid; u < 3 * VCP_L1_WIN_H * (PP_WIN_USED / 16); u += PP_THREADS) {
        const int k = 1 + u / (VCP_L1_WIN_H * (PP_WIN_USED / 16)), rem = u % (VCP_L1_WIN_H * (PP_WIN_USED / 16));
        const int r = rem / (PP_WIN_USED / 16), j = (rem % (PP_WIN_USED / 16)) * 16;
        const uint4 a = *reinterpret_cast<const uint4*>(&S.win[0][r][j]);
        const uint32_t e = *reinterpret_cast<const uint32_t*>(&S.win[0][r][j + 16]);
        const uint32_t sh = 8u * (uint32_t)k;
        *reinterpret_cast<uint4*>(&S.win[k][r][j]) = make_uint4(__funnelshift_r(a.x, a.y, sh), __funnelshift_r(a.y, a.z, sh),
                                                               __funnelshift_r(a.z, a.w, sh), __funnelshift_r(a.w, e, sh));
    }
    __syncthreads();

    // ---- L1: thread = (macroblock, dx) --------------------------------------------------------
    if (tid < PP_ITEMS) {
        const int grp = tid >> 4, mbl = tid & 15;
        // dx order: the two half-warps of a warp get word offsets of different parity, so that together they read 32
        // distinct banks: even groups take dx index 0-3, 8-11, 16-19, 24, odd groups 4-7, 12-15, 20-23
        const int gi2 = grp >> 1;
        const int dxi = (gi2 >> 2) * 8 + (gi2 & 3) + ((grp & 1) ? 4 : 0);
        if (mx0 + mbl < g.mbw) {
            const int col = 8 * mbl + dxi + 4;           // window column of this thread's candidates (window starts at -16)
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(&S.win[col & 3][0][col & ~3]);
            uint32_t cur[8][2];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const uint2 v = *reinterpret_cast<const uint2*>(&S.hcur[r][8 * mbl]);
                cur[r][0] = v.x; cur[r][1] = v.y;
            }
            // the vector cost is folded into the accumulators' initial values
            const uint32_t penx = VCP_ME_L1_PEN * (uint32_t)vcp_iabs(dxi - VCP_ME_R1);
            uint32_t acc[PP_NDX];
#pragma unroll
            for (int i = 0; i < PP_NDX; i++) acc[i] = penx + VCP_ME_L1_PEN * (uint32_t)(i < VCP_ME_R1 ? VCP_ME_R1 - i : i - VCP_ME_R1);
#pragma unroll
            for (int rr = 0; rr < VCP_L1_WIN_H; rr++) {
                const uint32_t r0 = wp[rr * (VCP_L1_WIN_W / 4)], r1 = wp[rr * (VCP_L1_WIN_W / 4) + 1];
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const int di = rr - r;  // dy + R1
                    if (di >= 0 && di < PP_NDX) acc[di] = sad4(r1, cur[r][1], sad4(r0, cur[r][0], acc[di]));
                }
            }
            uint32_t best = 0xffffffffu;
#pragma unroll
            for (int di = 0; di < PP_NDX; di++) {
                const uint32_t key = (acc[di] << 16) + (uint32_t)(di * PP_NDX) + (uint32_t)dxi;
                best = key < best ? key : best;
            }
            atomicMin(&S.best1[mbl], best);
        }
    }
    __syncthreads();

    if (tid < PP_MBS && mx0 + tid < g.mbw) {
        const int bi = (int)(S.best1[tid] & 0xffff);
        b.mvfp[(size_t)n * g.nmb + my * g.mbw + mx0 + tid] = make_short2((short)(2 * (bi % PP_NDX - VCP_ME_R1)), (short)(2 * (bi / PP_NDX - VCP_ME_R1)));
    }
}

// L0: warp = macroblock, +-2 around twice the L1 vector on the full-res originals
__global__ void __launch_bounds__(PP0_WARPS * 32) me_prepass_l0_kernel(VcpGeom g, VcpBufs b, int nframes, int gop, int t, int g0,
                                                                        const __grid_constant__ VcpTmaps tm) {
    __shared__ PrepassL0Warp sh[PP0_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = t < 0 ? (int)blockIdx.z : ((int)blockIdx.z + g0) * gop + t;
    if (n >= nframes || n % gop == 0) return;
    const int my = blockIdx.y, mx = blockIdx.x * PP0_WARPS + warp;
    if (mx >= g.mbw) return;
    PrepassL0Warp& S = sh[warp];
    if (lane == 0) { mbar_init(&S.bar, 1); mbar_init_fence(); }
    __syncwarp();
    const short2 l1 = b.mvfp[(size_t)n * g.nmb + my * g.mbw + mx];
    const int cx = l1.x, cy = l1.y;
    const int xr = 16 * mx + cx - 2 + VCP_PAD;       // plane column of the region's first sample
    if (lane == 0) {
        mbar_expect_tx(&S.bar, (uint32_t)(VCP_L0_REF_H * VCP_L0_REF_W + 256));
        tma_load_3d(&S.ref[0], &tm.y_ref, &S.bar, xr & ~15, 16 * my + cy - 2 + VCP_PAD, n - 1);
        tma_load_3d(&S.cur[0][0], &tm.y_cur, &S.bar, 16 * mx + VCP_PAD, 16 * my + VCP_PAD, n);
    }
    mbar_wait(&S.bar, 0);
    uint32_t acc[5] = {0, 0, 0, 0, 0};
    const int ox = lane % 5, rg = lane / 5;   // lanes 0..19: horizontal offset, group of four current rows
    if (lane < 20) {
        uint4 c[4];
#pragma unroll
        for (int i = 0; i < 4; i++) c[i] = *reinterpret_cast<const uint4*>(&S.cur[4 * rg + i][0]);
        const int o = (xr & 15) + ox;                                // byte of this lane's first column inside a region row
        const uint32_t* rp = reinterpret_cast<const uint32_t*>(&S.ref[(4 * rg) * VCP_L0_REF_W]) + (o >> 2);
        const uint32_t sh = (uint32_t)(o & 3) * 8u;
#pragma unroll
        for (int tt = 0; tt < 8; tt++) {
            const uint32_t* q = rp + tt * (VCP_L0_REF_W / 4);
            const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
            const uint32_t r0 = __funnelshift_r(w0, w1, sh), r1 = __funnelshift_r(w1, w2, sh), r2 = __funnelshift_r(w2, w3, sh), r3 = __funnelshift_r(w3, w4, sh);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int oy = tt - i;
                if (oy >= 0 && oy < 5)
                    acc[oy] = sad4(r3, c[i].w, sad4(r2, c[i].z, sad4(r1, c[i].y, sad4(r0, c[i].x, acc[oy]))));
            }
        }
    }
    uint32_t best = 0xffffffffu;
#pragma unroll
    for (int oy = 0; oy < 5; oy++) {
        uint32_t v = acc[oy];
        v += __shfl_down_sync(0xffffffffu, v, 10);
        v += __shfl_down_sync(0xffffffffu, v, 5);
        if (lane < 5) {
            const int mvx = cx + ox - 2, mvy = cy + oy - 2;
            const uint32_t key = ((v + VCP_ME_L0_PEN * (uint32_t)(vcp_iabs(mvx) + vcp_iabs(mvy))) << 8) | (uint32_t)(oy * 5 + ox);
            best = key < best ? key : best;
        }
    }
    best = warp_min(best);
    if (lane == 0) {
        const int k = (int)(best & 0xff);
        b.mvfp[(size_t)n * g.nmb + my * g.mbw + mx] = make_short2((short)(cx + k % 5 - 2), (short)(cy + k / 5 - 2));
    }
}

// predictor estimate from the neighbours' pre-pass vectors (oracle: pmv_estimate)
__device__ __forceinline__ void pmv_estimate(const VcpGeom& g, const VcpBufs& b, const short2* __restrict__ mvfp, int mx, int my, int& px, int& py) {
    const int row0 = vcp_row_first(b, my);
    const bool aA = mx > 0, aB = my > row0, aC = aB && mx + 1 < g.mbw, aD = aB && mx > 0;
    const int i = my * g.mbw + mx;
    int ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0;
    if (aA) { short2 v = mvfp[i - 1]; ax = 4 * v.x; ay = 4 * v.y; }
    if (aB) { short2 v = mvfp[i - g.mbw]; bx = 4 * v.x; by = 4 * v.y; }
    if (aC) { short2 v = mvfp[i - g.mbw + 1]; cx = 4 * v.x; cy = 4 * v.y; }
    else if (aD) { short2 v = mvfp[i - g.mbw - 1]; cx = 4 * v.x; cy = 4 * v.y; }
    if (!aB && aA) { px = ax; py = ay; return; }
    px = vcp_median3(ax, bx, cx); py = vcp_median3(ay, by, cy);
}

// ---------------------------------------------------------------------------------------------
// K2b
// ---------------------------------------------------------------------------------------------
constexpr int RF_WARPS = 4;
constexpr int RF_ROWW = 6;                               // words per row of the working copy (24 B: conflict-free for 16 rows)
constexpr int RF_PLANEW = VCP_RF_WIN_H * RF_ROWW;        // 120 words per plane

// Per warp: where the copy engine lands the boxes (48-byte rows starting on the 16-byte boundary below the
// window, as the tensor map dictates) and the working copy the SAD loops read (24-byte rows with the
// misalignment removed: the 16 rows of a macroblock fall into distinct banks and the byte shifts of the
// full-pel candidates are compile-time).  In the working copy, sample (x, y) relative to the macroblock
// displaced by the window's vector sits at byte x + 2 of row y + 2.
constexpr int RF_LANDW = VCP_RF_WIN_W / 4;                // 12 words per landed row
constexpr int RF_LANDPLANE = VCP_RF_WIN_H * RF_LANDW;     // 240 words per landed plane
struct __align__(128) RefineWarp {
    uint32_t land[4][VCP_RF_WIN_H][RF_LANDW];
    uint32_t landz[256];      // VCP_RF_WIN_H x RF_LANDW, padded to a multiple of 128 bytes
    uint32_t landp[256];
    uint32_t w[4][VCP_RF_WIN_H][RF_ROWW];
    uint64_t bar;
    uint64_t barq;            // HEVC quarter-sample step: the 24 x 24 window of integer samples has landed in landz / landp
};
static_assert(VCP_RFQ_WIN_H * VCP_RFQ_WIN_W <= 2 * 256 * 4, "the quarter-sample window lands in landz + landp");
static_assert(HQ_WROWS * HQ_WPITCH <= 4 * RF_LANDPLANE, "the row-pass output of the quarter-sample step reuses the landing planes");
static_assert(RF_LANDPLANE <= 256, "landing buffers");

// landing buffer -> working copy, planes [p0, p1), dropping the window's misalignment `mis` (0..15):
// 60 eight-byte units per plane, two per lane
__device__ __forceinline__ void rf_restage(RefineWarp& S, int p0, int p1, int lane, int mis) {
    const int u0 = lane, u1 = lane + 32;
    const int r0 = u0 / 3, j0 = u0 % 3, r1 = u1 / 3, j1 = u1 % 3;
    const uint32_t sh = (uint32_t)(mis & 3) * 8u;
    const uint32_t* q0 = &S.land[0][r0][2 * j0 + (mis >> 2)];
    const uint32_t* q1 = &S.land[0][r1 < VCP_RF_WIN_H ? r1 : 0][2 * j1 + (mis >> 2)];
    for (int p = p0; p < p1; p++) {
        {
            const uint32_t a0 = q0[p * RF_LANDPLANE], a1 = q0[p * RF_LANDPLANE + 1], a2 = q0[p * RF_LANDPLANE + 2];
            *reinterpret_cast<uint2*>(&S.w[p][r0][2 * j0]) = make_uint2(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh));
        }
        if (u1 < 60) {
            const uint32_t a0 = q1[p * RF_LANDPLANE], a1 = q1[p * RF_LANDPLANE + 1], a2 = q1[p * RF_LANDPLANE + 2];
            *reinterpret_cast<uint2*>(&S.w[p][r1][2 * j1]) = make_uint2(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh));
        }
    }
}

// SAD of this lane's 8 samples against the macroblock at the ORIGIN vector of a landed G window with misalignment `mis`
__device__ __forceinline__ int rf_sad_land(const uint32_t* L, int row, int hx, int mis, uint2 c8) {
    const int o = mis + 2 + hx;                                        // byte of this lane's first sample inside a landed row
    const uint32_t* r = L + (row + 2) * RF_LANDW + (o >> 2);
    const uint32_t sh = (uint32_t)(o & 3) * 8u;
    const uint32_t a0 = r[0], a1 = r[1], a2 = r[2];
    return warp_sum((int)sad4(__funnelshift_r(a1, a2, sh), c8.y, sad4(__funnelshift_r(a0, a1, sh), c8.x, 0)));
}

// HQ: HEVC quarter-sample candidates ranked by their exact prediction (hevc_subpel = 3): a kernel of its own, so that the
// H.264 / proxy-ranked build keeps its registers
template <bool HQ>
__global__ void __launch_bounds__(RF_WARPS * 32) me_refine_kernel(VcpGeom g, VcpBufs b, VcpStep s, const __grid_constant__ VcpTmaps tm) {
    __shared__ RefineWarp sh[RF_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * RF_WARPS + warp;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    RefineWarp& S = sh[warp];
    if (lane == 0) { mbar_init(&S.bar, 1); if (HQ) mbar_init(&S.barq, 1); mbar_init_fence(); }
    __syncwarp();
    const int n = vcp_frame_of(s, gi);
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    const int qp = b.qp[n];
    const int lam = vcp_lambda(qp);
    const short2* mvfp = b.mvfp + (size_t)n * g.nmb;
    int pmx, pmy;
    pmv_estimate(g, b, mvfp, mx, my, pmx, pmy);
    const short2 f = mvfp[mbi];
    const int pvx = vcp_clip3(-VCP_MV_FP_MAX, VCP_MV_FP_MAX, (pmx + 2) >> 2);
    const int pvy = vcp_clip3(-VCP_MV_FP_MAX, VCP_MV_FP_MAX, (pmy + 2) >> 2);
    const int slot = vcp_rec_slot(s, gi, s.t - 1);
    const bool sub = !g.hevc || g.hevc_subpel;   // the half-sample planes exist and are searched
    // The zero vector and the rounded predictor are candidates 9 and 10.  When one of them coincides with an
    // earlier candidate (the pre-pass vector or its 8 neighbours; the predictor also with the zero vector) it has
    // that candidate's cost and a larger index, so it can never win: it is not evaluated and needs no window.
    const bool evalZ = vcp_iabs(f.x) > 1 || vcp_iabs(f.y) > 1;
    const bool evalP = (vcp_iabs(pvx - f.x) > 1 || vcp_iabs(pvy - f.y) > 1) && (pvx | pvy) != 0;
    const int x0 = VCP_PAD + 16 * mx - 2, y0 = VCP_PAD + 16 * my - 2;   // plane coordinates of window sample (-2,-2) at vector 0
    if (lane == 0) {
        const uint32_t box = VCP_RF_WIN_H * VCP_RF_WIN_W;
        mbar_expect_tx(&S.bar, (sub ? 4u : 1u) * box + (evalZ ? box : 0u) + (evalP ? box : 0u));
        // boxes start on the 16-byte boundary below the window (x0 is 14 mod 16: the misalignment of a window is (14 + vx) & 15)
        tma_load_4d(&S.land[0][0][0], sub ? &tm.rec4 : &tm.rec1, &S.bar, (x0 + f.x) & ~15, y0 + f.y, 0, slot);
        if (evalZ) tma_load_4d(&S.landz[0], &tm.rec1, &S.bar, x0 & ~15, y0, 0, slot);
        if (evalP) tma_load_4d(&S.landp[0], &tm.rec1, &S.bar, (x0 + pvx) & ~15, y0 + pvy, 0, slot);
    }
    const uint8_t* yc = b.src_y + (size_t)n * g.ysize + g.yoff;
    const int row = lane >> 1, hx = (lane & 1) * 8;
    const int px = 16 * mx, py = 16 * my;
    const uint2 c8 = *reinterpret_cast<const uint2*>(yc + (size_t)(py + row) * g.ys + px + hx);

    // Scalar work (candidate vectors, vector costs) is done once by the lane whose index equals the candidate;
    // the 32 lanes share only the SAD loop.
    int cvx, cvy;   // candidate `lane`
    {
        const int k = lane;
        if (k == 0) { cvx = f.x; cvy = f.y; }
        else if (k < 9) { const int q = k - 1 + (k > 4); cvx = f.x + q % 3 - 1; cvy = f.y + q / 3 - 1; }
        else if (k == 9) { cvx = 0; cvy = 0; }
        else { cvx = pvx; cvy = pvy; }
    }
    int mycost = lam * (vcp_se_len(4 * cvx - pmx) + vcp_se_len(4 * cvy - pmy));

    mbar_wait(&S.bar, 0);
    rf_restage(S, 0, 1, lane, (x0 + f.x) & 15);
    __syncwarp();

    // full-pel: the pre-pass vector and its 8 neighbours, three rows of the G plane, three byte shifts each
    {
        const uint32_t* wl = &S.w[0][row + 1][hx >> 2];   // row of dy = -1, bytes hx .. hx+11
#pragma unroll
        for (int dyi = 0; dyi < 3; dyi++) {
            const uint2 a = *reinterpret_cast<const uint2*>(wl + dyi * RF_ROWW);
            const uint32_t c = wl[dyi * RF_ROWW + 2];
#pragma unroll
            for (int dxi = 0; dxi < 3; dxi++) {
                const uint32_t shb = 8u * (dxi + 1);
                const int sad = warp_sum((int)sad4(__funnelshift_r(a.y, c, shb), c8.y, sad4(__funnelshift_r(a.x, a.y, shb), c8.x, 0)));
                const int q = dyi * 3 + dxi, k = q == 4 ? 0 : (q < 4 ? q + 1 : q);
                if (lane == k) mycost += sad;
            }
        }
    }
    constexpr int NEVER = 0x07ffffff;
    if (evalZ) { const int s9 = rf_sad_land(S.landz, row, hx, x0 & 15, c8); if (lane == 9) mycost += s9; }
    else if (lane == 9) mycost = NEVER;
    if (evalP) { const int s10 = rf_sad_land(S.landp, row, hx, (x0 + pvx) & 15, c8); if (lane == 10) mycost += s10; }
    else if (lane == 10) mycost = NEVER;
    uint32_t best = warp_min(lane < 11 ? (((uint32_t)mycost << 4) | (uint32_t)lane) : 0xffffffffu);
    const int kb = (int)(best & 15);
    const int bvx = __shfl_sync(0xffffffffu, cvx, kb), bvy = __shfl_sync(0xffffffffu, cvy, kb);
    uint32_t bcost = best >> 4;
    if (bcost < VCP_SUBPEL_SKIP_COST) {   // warp-uniform
        if (lane == 0) {
            b.mv[(size_t)gi * g.nmb + mbi] = make_short2((short)(4 * bvx), (short)(4 * bvy));
            b.mbtype[(size_t)gi * g.nmb + mbi] = VCP_MB_P16;
        }
        return;
    }

    int ox = 0, oy = 0;
    if (HQ) {
        // the integer samples -4 .. 19 around the block at its best full-sample vector, for the quarter-sample step: the
        // zero-vector / predictor landing buffers have been read (rf_sad_land above) and take the box
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            mbar_expect_tx(&S.barq, VCP_RFQ_WIN_H * VCP_RFQ_WIN_W);
            tma_load_4d(&S.landz[0], &tm.rec1q, &S.barq, (x0 + bvx - 2) & ~15, y0 + bvy - 2, 0, slot);
        }
    }
    if (sub) {   // HEVC: full samples only, or (hevc_subpel) the half-sample step alone
        int bdx, bdy;   // best full-pel position relative to the vector the working copy is centred on
        if (kb >= 9) {
            // the zero vector or the predictor won and lies outside the window: bring the four planes around it
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(&S.bar, 4u * VCP_RF_WIN_H * VCP_RF_WIN_W);
                tma_load_4d(&S.land[0][0][0], &tm.rec4, &S.bar, (x0 + bvx) & ~15, y0 + bvy, 0, slot);
            }
            mbar_wait(&S.bar, 1);
            rf_restage(S, 0, 4, lane, (x0 + bvx) & 15);
            bdx = 0; bdy = 0;
        } else {
            rf_restage(S, 1, 4, lane, (x0 + f.x) & 15);
            bdx = bvx - f.x; bdy = bvy - f.y;
        }
        __syncwarp();
        const uint32_t* Ww = &S.w[0][0][0];
        const int lb = (row + 2 + bdy) * (4 * RF_ROWW) + hx + 2 + bdx;   // byte of this lane's first sample at the best full-pel position, plane G

        // ---- half-sample step: the 8 neighbours are fixed samples of the planes b (1), h (2), j (3) ----
        {
            int cq, cr;
            { const int k = (lane - 1) & 7, q = k + (k > 3); cq = (q % 3 - 1) * 2; cr = (q / 3 - 1) * 2; }
            mycost = lam * (vcp_se_len(4 * bvx + cq - pmx) + vcp_se_len(4 * bvy + cr - pmy));
            if (lane == 0) mycost = (int)bcost;
            // samples one column to the left (ix = -1) and at the position (ix = 0) come out of the same three words
            const uint32_t* wm = Ww + ((lb - 1) >> 2);
            const uint32_t sm = (uint32_t)((lb - 1) & 3) * 8u;
#define RF_PAIR(PLANE, IY, KL, KR)                                                                                      \
            {                                                                                                           \
                const uint32_t* r_ = wm + (PLANE) * RF_PLANEW + (IY) * RF_ROWW;                                          \
                const uint32_t a0 = r_[0], a1 = r_[1], a2 = r_[2];                                                       \
                if ((KL) > 0) {                                                                                         \
                    const int sad = warp_sum((int)sad4(__funnelshift_r(a1, a2, sm), c8.y, sad4(__funnelshift_r(a0, a1, sm), c8.x, 0))); \
                    if (lane == (KL)) mycost += sad;                                                                    \
                }                                                                                                       \
                {                                                                                                       \
                    const int sad = warp_sum((int)sad4(__funnelshift_rc(a1, a2, sm + 8u), c8.y, sad4(__funnelshift_rc(a0, a1, sm + 8u), c8.x, 0))); \
                    if (lane == (KR)) mycost += sad;                                                                    \
                }                                                                                                       \
            }
            RF_PAIR(3, -1, 1, 3)   // j above: (-2,-2), (2,-2)
            RF_PAIR(2, -1, 0, 2)   // h above: (0,-2)
            RF_PAIR(1, 0, 4, 5)    // b: (-2,0), (2,0)
            RF_PAIR(3, 0, 6, 8)    // j below: (-2,2), (2,2)
            RF_PAIR(2, 0, 0, 7)    // h below: (0,2)
#undef RF_PAIR
            best = warp_min(lane < 9 ? (((uint32_t)mycost << 4) | (uint32_t)lane) : 0xffffffffu);
            bcost = best >> 4;
            const int bk = (int)(best & 15);
            ox = __shfl_sync(0xffffffffu, bk ? cq : 0, bk);
            oy = __shfl_sync(0xffffffffu, bk ? cr : 0, bk);
        }
        // ---- quarter-sample step (H.264: not in the fast -preset tiers.  HEVC with hevc_subpel = 2: its quarter positions
        //      are NOT averages of half-sample planes, but the averages rank the eight candidates as well as the exact
        //      7/8-tap predictions do (oracle: hevc_luma_proxy; measured on quarter-sample pans 53 591 vs 53 577 bytes), and
        //      hevc_p_recon codes the exact prediction of whichever vector wins.  hevc_subpel = 3 ranks by the exact
        //      prediction: me_refine_kernel<true>, below) ----
        if (g.hevc ? g.hevc_subpel == 2 : g.effort > 0) {
            // The eight quarter positions around the best half position c are averages of samples of the 3 x 3 half-sample
            // neighbourhood H[i][j] of c (i, j in -1..1): horizontal / vertical neighbours average H[0][0] with H[0][dx] /
            // H[dy][0]; the diagonal ones average H[0][dx] with H[dy][0] when c's two half coordinates have the same parity
            // (c is an integer or a j sample) and H[0][0] with H[dy][dx] otherwise (8.4.2.2.1: the pair "x odd, y even" +
            // "x even, y odd").  Nine fetches instead of sixteen; lane m < 9 works out where H[m/3-1][m%3-1] lives (every
            // lane's first sample has the same misalignment: rows are 24 B, halves 8 B apart) and broadcasts it packed as
            // word * 32 + shift.
            const int hx0 = ox >> 1, hy0 = oy >> 1;                      // c in half-sample units: -1, 0, 1
            const int lbr = lb & 3;
            const uint32_t* Wl = Ww + (lb >> 2);
            int pk = 0;
            {
                const int m = lane < 9 ? lane : 0;
                const int X = hx0 + m % 3 - 1, Y = hy0 + m / 3 - 1;
                const int a = lbr + (hpel_plane(X, Y) * VCP_RF_WIN_H + (Y >> 1)) * (4 * RF_ROWW) + (X >> 1);
                pk = (a >> 2) * 32 + (a & 3) * 8;
            }
            uint2 H[9];
#pragma unroll
            for (int m = 0; m < 9; m++) {
                const int p = __shfl_sync(0xffffffffu, pk, m);
                const uint32_t* r = Wl + (p >> 5);
                const uint32_t shb = (uint32_t)p & 31u;
                const uint32_t w0 = r[0], w1 = r[1], w2 = r[2];
                H[m] = make_uint2(__funnelshift_r(w0, w1, shb), __funnelshift_r(w1, w2, shb));
            }
            int cq, cr;
            { const int k = (lane - 1) & 7, q = k + (k > 3); cq = ox + (q % 3 - 1); cr = oy + (q / 3 - 1); }
            mycost = lam * (vcp_se_len(4 * bvx + cq - pmx) + vcp_se_len(4 * bvy + cr - pmy));
            if (lane == 0) mycost = (int)bcost;
            const bool same_par = ((hx0 ^ hy0) & 1) == 0;
#pragma unroll
            for (int k = 1; k <= 8; k++) {
                const int q = (k - 1) + (k - 1 > 3), dx = q % 3 - 1, dy = q / 3 - 1;   // compile-time
                uint2 p1, p2;
                if (dy == 0) { p1 = H[4]; p2 = H[4 + dx]; }
                else if (dx == 0) { p1 = H[4]; p2 = H[4 + 3 * dy]; }
                else {
                    p1 = same_par ? H[4 + dx] : H[4];
                    p2 = same_par ? H[4 + 3 * dy] : H[4 + 3 * dy + dx];
                }
                const int sad = warp_sum((int)sad4(vcp_avg4(p1.y, p2.y), c8.y, sad4(vcp_avg4(p1.x, p2.x), c8.x, 0)));
                if (lane == k) mycost += sad;
            }
            best = warp_min(lane < 9 ? (((uint32_t)mycost << 4) | (uint32_t)lane) : 0xffffffffu);
            bcost = best >> 4;
            const int bk = (int)(best & 15);
            ox = __shfl_sync(0xffffffffu, bk ? cq : ox, bk);
            oy = __shfl_sync(0xffffffffu, bk ? cr : oy, bk);
        }
    }
    if (HQ) {
        // ---- HEVC quarter-sample step, exact (hevc_subpel = 3; oracle: hevc_refine_cu, step 1): the 8 neighbours of the best
        //      half-sample position, each predicted exactly as the decoder will (vcp_hevc_qpel.cuh).  Candidates in one
        //      column of the 3 x 3 ring share their row pass.
        mbar_wait(&S.barq, 0);
        __syncwarp();                                   // every lane is done with the half-sample planes: W takes their place
        uint32_t* W = &S.land[0][0][0];
        const uint32_t* win = &S.landz[0];
        const int misq = (x0 + bvx - 2) & 15;
        int cq, cr;
        { const int k = (lane - 1) & 7, q = k + (k > 3); cq = ox + (q % 3 - 1); cr = oy + (q / 3 - 1); }
        mycost = lam * (vcp_se_len(4 * bvx + cq - pmx) + vcp_se_len(4 * bvy + cr - pmy));
        if (lane == 0) mycost = (int)bcost;
#pragma unroll 1
        for (int dxi = 0; dxi < 3; dxi++) {
            const int qx = ox + dxi - 1;
            hq_hpass(win, VCP_RFQ_WIN_W / 4, misq + (qx >> 2) + 1, qx & 3, W, lane);
            __syncwarp();
#pragma unroll 1
            for (int dyi = 0; dyi < 3; dyi++) {
                if (dxi == 1 && dyi == 1) continue;
                const int qy = oy + dyi - 1;
                const uint2 p8 = hq_vpass(W, row + (qy >> 2) + 1, lane & 1, qy & 3);
                const int sad = warp_sum((int)sad4(p8.y, c8.y, sad4(p8.x, c8.x, 0)));
                const int q = dyi * 3 + dxi, k = q < 4 ? q + 1 : q;
                if (lane == k) mycost += sad;
            }
            __syncwarp();
        }
        best = warp_min(lane < 9 ? (((uint32_t)mycost << 4) | (uint32_t)lane) : 0xffffffffu);
        bcost = best >> 4;
        const int bk = (int)(best & 15);
        ox = __shfl_sync(0xffffffffu, bk ? cq : ox, bk);
        oy = __shfl_sync(0xffffffffu, bk ? cr : oy, bk);
    }
    // intra or inter?  Intra16x16 estimated on the ORIGINAL picture (best of V / H / DC from original
    // neighbours): no reconstruction needed, so the decision stays macroblock-parallel (oracle:
    // intra_estimate, vcp_intra_wins).  Intra macroblocks are coded by i_fix_kernel after the inter ones.
    int type = VCP_MB_P16;
    {
        const int row0 = vcp_row_first(b, my);
        const bool aL = mx > 0, aT = my > row0;
        const uint8_t* cr = yc + (size_t)(py + row) * g.ys + px;
        const uint2 top8 = *reinterpret_cast<const uint2*>(yc + (ptrdiff_t)(py - 1) * g.ys + px + hx);
        const uint32_t left = cr[-1];
        const uint32_t l4 = left * 0x01010101u;
        const int st = warp_sum(lane < 2 ? (int)sad4(top8.x, 0u, sad4(top8.y, 0u, 0u)) : 0);   // 16 samples above
        const int sl = warp_sum((lane & 1) ? 0 : (int)left);                                     // 16 samples to the left
        const int dc = (aT && aL) ? (st + sl + 16) >> 5 : aT ? (st + 8) >> 4 : aL ? (sl + 8) >> 4 : 128;
        const uint32_t d4 = (uint32_t)dc * 0x01010101u;
        const int sv = warp_sum((int)sad4(c8.y, top8.y, sad4(c8.x, top8.x, 0u)));
        const int shh = warp_sum((int)sad4(c8.y, l4, sad4(c8.x, l4, 0u)));
        const int sd = warp_sum((int)sad4(c8.y, d4, sad4(c8.x, d4, 0u)));
        int best_i = sd;
        if (aT && sv < best_i) best_i = sv;
        if (aL && shh < best_i) best_i = shh;
        if (vcp_intra_wins(best_i, (int)bcost, lam)) type = VCP_MB_I16;
    }
    if (lane == 0) {
        b.mv[(size_t)gi * g.nmb + mbi] = make_short2((short)(4 * bvx + ox), (short)(4 * bvy + oy));
        b.mbtype[(size_t)gi * g.nmb + mbi] = (uint8_t)type;
        if (type == VCP_MB_I16) atomicAdd(&b.icount[(size_t)gi * g.slices + vcp_row_slice(b, my)], 1);
    }
}

}  // namespace

// t < 0: every resident frame.  t >= 0: picture t of GOPs [g0, g1) (g1 < 0: up to the last GOP)
void vcp_launch_me_prepass(const VcpGeom& g, const VcpBufs& b, const VcpTmaps& tm, int nframes, int gop, int t, cudaStream_t st, int g0, int g1) {
    if (nframes <= 0) return;
    int nz = t < 0 ? nframes : (nframes - t + gop - 1) / gop;   // GOPs that own a picture t
    if (t >= 0) { if (g1 >= 0 && g1 < nz) nz = g1; nz -= g0; } else g0 = 0;
    if (nz <= 0) return;
    dim3 grid((g.mbw + PP_MBS - 1) / PP_MBS, g.mbh, nz);
    me_prepass_kernel<<<grid, PP_THREADS, 0, st>>>(g, b, nframes, gop, t, g0, tm);
    dim3 grid0((g.mbw + PP0_WARPS - 1) / PP0_WARPS, g.mbh, nz);
    me_prepass_l0_kernel<<<grid0, PP0_WARPS * 32, 0, st>>>(g, b, nframes, gop, t, g0, tm);
}

void vcp_launch_me_refine(const VcpGeom& g, const VcpBufs& b, const VcpTmaps& tm, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + RF_WARPS - 1) / RF_WARPS, s.ngop);
    if (g.hevc && g.hevc_subpel >= 3) me_refine_kernel<true><<<grid, RF_WARPS * 32, 0, st>>>(g, b, s, tm);
    else me_refine_kernel<false><<<grid, RF_WARPS * 32, 0, st>>>(g, b, s, tm);
}
