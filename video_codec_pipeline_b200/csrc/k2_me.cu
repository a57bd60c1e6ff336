// K2 — integer motion estimation.
//
//  K2a  me_prepass_kernel : hierarchical full-pel search on the ORIGINAL frames.  It does not
//        depend on any reconstruction, so one launch covers every P-frame of every GOP that
//        is resident — this is where the bulk of the SAD work lives and it is off the
//        frame-to-frame dependency chain.
//          L1: half-res 8x8 block per macroblock, exhaustive +-12 (625 candidates), SAD via
//              VABSDIFF4.U8.ACC (__vsadu4); lanes own a column offset dx, walk the 32
//              reference rows once and feed 8 running accumulators (one per dy in flight).
//          L0: full-res 16x16, +-2 around twice the L1 vector, warp-cooperative SAD +
//              redux.sync reduction.
//  K2b  me_refine_kernel  : inside the per-frame chain, against the RECONSTRUCTED reference:
//        11 full-pel candidates, then 8 half-pel and 8 quarter-pel neighbours evaluated from
//        half-pel planes (b, h, j of 8.4.2.2.1) built once per macroblock in shared memory.
//
// Replaces x264's `me=hex subme=7` / NVENC's ME inside the ffmpeg child
// (/root/reference/cmd/consumer.go:376-382).  Decisions are bit-identical to
// oracle/h264_oracle.c (me_prepass, me_refine_mb).
#include "vcp_dev.cuh"
#include "vcp_luma_interp.cuh"

namespace {

constexpr int ME_WARPS = 8;                       // macroblocks per CTA (one row segment)
constexpr int L1_WIN_X0 = 16;                     // window starts 16 px left of the CTA's first block (16 B aligned)
constexpr int L1_WIN_WA = 112;                    // 16 + 64 + 12 + slack for unaligned 12-byte reads, 7 x 16 B
constexpr int L1_WIN_H = 8 + 2 * VCP_ME_R1;       // 32
constexpr int L0_W = 28;                          // 16 + 4 candidates + slack, 7 words

struct __align__(16) PrepassWarp {
    uint8_t cur[16][16];
    uint8_t ref[20][L0_W];
};

// grid z = frame (t < 0: every resident frame) or GOP (t >= 0: picture t of every GOP, so that the
// pre-pass of later pictures runs beside the reconstruction chain of earlier ones)
__global__ void __launch_bounds__(ME_WARPS * 32) me_prepass_kernel(VcpGeom g, VcpBufs b, int nframes, int gop, int t, int g0) {
    __shared__ __align__(16) uint8_t win[L1_WIN_H][L1_WIN_WA];
    __shared__ PrepassWarp pw[ME_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = t < 0 ? (int)blockIdx.z : ((int)blockIdx.z + g0) * gop + t;
    if (n >= nframes || n % gop == 0) return;  // IDR: no search
    const int my = blockIdx.y, mx0 = blockIdx.x * ME_WARPS, mx = mx0 + warp;
    const uint8_t* hc = b.src_h + (size_t)n * g.hsize + g.hoff;
    const uint8_t* hp = b.src_h + (size_t)(n - 1) * g.hsize + g.hoff;
    // stage the shared L1 window with 16-byte loads: rows 8my-12 .. 8my+19, cols 8mx0-16 .. 8mx0+95
    {
        const uint8_t* base = hp + (ptrdiff_t)(8 * my - VCP_ME_R1) * g.hs + (8 * mx0 - L1_WIN_X0);
        for (int i = threadIdx.x; i < L1_WIN_H * (L1_WIN_WA / 16); i += blockDim.x) {
            const int r = i / (L1_WIN_WA / 16), c = i % (L1_WIN_WA / 16);
            reinterpret_cast<uint4*>(&win[r][0])[c] = __ldg(reinterpret_cast<const uint4*>(base + (ptrdiff_t)r * g.hs) + c);
        }
    }
    __syncthreads();
    if (mx >= g.mbw) return;

    // ---- L1 -----------------------------------------------------------------------------
    uint32_t cur[8][2];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint2 v = *reinterpret_cast<const uint2*>(hc + (size_t)(8 * my + r) * g.hs + 8 * mx);
        cur[r][0] = v.x; cur[r][1] = v.y;
    }
    uint32_t best = 0xffffffffu;
    if (lane < 2 * VCP_ME_R1 + 1) {
        const int dx = lane - VCP_ME_R1;
        const int col = 8 * warp + lane + (L1_WIN_X0 - VCP_ME_R1);  // window column of this lane's candidates
        uint32_t acc[2 * VCP_ME_R1 + 1];
#pragma unroll
        for (int i = 0; i < 2 * VCP_ME_R1 + 1; i++) acc[i] = 0;
#pragma unroll
        for (int rr = 0; rr < L1_WIN_H; rr++) {
            const uint2 ref = ld8_unaligned(&win[rr][col]);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int di = rr - r;  // dy + R1
                if (di >= 0 && di <= 2 * VCP_ME_R1)
                    acc[di] = sad4(ref.y, cur[r][1], sad4(ref.x, cur[r][0], acc[di]));
            }
        }
#pragma unroll
        for (int di = 0; di < 2 * VCP_ME_R1 + 1; di++) {
            const int dy = di - VCP_ME_R1;
            const uint32_t cost = acc[di] + VCP_ME_L1_PEN * (vcp_iabs(dx) + vcp_iabs(dy));
            const uint32_t key = (cost << 16) | (uint32_t)(di * (2 * VCP_ME_R1 + 1) + lane);
            best = key < best ? key : best;
        }
    }
    best = warp_min(best);
    const int bi = (int)(best & 0xffff), W = 2 * VCP_ME_R1 + 1;
    const int cx = 2 * (bi % W - VCP_ME_R1), cy = 2 * (bi / W - VCP_ME_R1);

    // ---- L0: +-2 around (cx,cy) on the full-res originals, one lane per candidate ------------
    const uint8_t* yc = b.src_y + (size_t)n * g.ysize + g.yoff;
    const uint8_t* yp = b.src_y + (size_t)(n - 1) * g.ysize + g.yoff;
    PrepassWarp& S = pw[warp];
    {
        const int row = lane >> 1, hx = (lane & 1) * 8;
        *reinterpret_cast<uint2*>(&S.cur[row][hx]) =
            *reinterpret_cast<const uint2*>(yc + (size_t)(16 * my + row) * g.ys + 16 * mx + hx);
        // reference region: rows cy-2 .. cy+17, cols cx-2 .. cx+25
        const uint8_t* rb = yp + (ptrdiff_t)(16 * my + cy - 2) * g.ys + 16 * mx + cx - 2;
        const int c = lane & 7, r0 = lane >> 3;   // 7 words per row, 4 rows per pass
        if (c < 7) {
#pragma unroll
            for (int r = r0; r < 20; r += 4)
                reinterpret_cast<uint32_t*>(&S.ref[r][0])[c] = ld4_unaligned(rb + (ptrdiff_t)r * g.ys + 4 * c);
        }
    }
    __syncwarp();
    best = 0xffffffffu;
    if (lane < 25) {
        const int ox = lane % 5, oy = lane / 5;   // candidate offset + 2
        uint32_t sad = 0;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const uint4 c4 = *reinterpret_cast<const uint4*>(&S.cur[r][0]);
            const uint32_t* q = reinterpret_cast<const uint32_t*>(&S.ref[r + oy][ox & ~3]);
            const uint32_t sh = (uint32_t)(ox & 3) * 8;
            const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
            sad = sad4(__funnelshift_r(w0, w1, sh), c4.x, sad);
            sad = sad4(__funnelshift_r(w1, w2, sh), c4.y, sad);
            sad = sad4(__funnelshift_r(w2, w3, sh), c4.z, sad);
            sad = sad4(__funnelshift_r(w3, w4, sh), c4.w, sad);
        }
        const int mvx = cx + ox - 2, mvy = cy + oy - 2;
        best = ((sad + VCP_ME_L0_PEN * (uint32_t)(vcp_iabs(mvx) + vcp_iabs(mvy))) << 8) | (uint32_t)lane;
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

constexpr int RF_WARPS = 4;

__global__ void __launch_bounds__(RF_WARPS * 32) me_refine_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ HpelWindow win[RF_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * RF_WARPS + warp;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    const int n = vcp_frame_of(s, gi);
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    const int qp = b.qp[n];
    const int lam = vcp_lambda(qp);
    const short2* mvfp = b.mvfp + (size_t)n * g.nmb;
    int pmx, pmy;
    pmv_estimate(g, b, mvfp, mx, my, pmx, pmy);
    const uint8_t* yc = b.src_y + (size_t)n * g.ysize + g.yoff;
    const uint8_t* yr = vcp_rec_luma(b, g, vcp_rec_slot(s, gi, s.t - 1)) + g.yoff;
    const int row = lane >> 1, hx = (lane & 1) * 8;
    const int px = 16 * mx, py = 16 * my;
    const uint2 c8 = *reinterpret_cast<const uint2*>(yc + (size_t)(py + row) * g.ys + px + hx);
    const short2 f = mvfp[mbi];
    const uint8_t* mb0 = yr + (ptrdiff_t)py * g.ys + px;               // sample (0,0) at mv (0,0)
    HpelWindow& W = win[warp];

    // Scalar work (candidate vectors, vector costs, plane addressing) is done once by the lane
    // whose index equals the candidate; the 32 lanes share only the SAD loop.
    const uint32_t* Ww = &W.w[0][0][0];

    // full-pel candidates: the pre-pass vector and its 8 neighbours from one staged window, plus
    // the zero vector and the rounded predictor from global memory
    int mis = hpel_window_stage(W, mb0 + (ptrdiff_t)f.y * g.ys + f.x, g.ys, g.ysize, lane, 1);
    const int pvx = vcp_clip3(-VCP_MV_FP_MAX, VCP_MV_FP_MAX, (pmx + 2) >> 2);
    const int pvy = vcp_clip3(-VCP_MV_FP_MAX, VCP_MV_FP_MAX, (pmy + 2) >> 2);
    int cvx, cvy;   // candidate `lane`
    {
        const int k = lane;
        if (k == 0) { cvx = f.x; cvy = f.y; }
        else if (k < 9) { const int q = k - 1 + (k > 4); cvx = f.x + q % 3 - 1; cvy = f.y + q / 3 - 1; }
        else if (k == 9) { cvx = 0; cvy = 0; }
        else { cvx = pvx; cvy = pvy; }
    }
    int mycost = lam * (vcp_se_len(4 * cvx - pmx) + vcp_se_len(4 * cvy - pmy));
#pragma unroll
    for (int k = 0; k < 9; k++) {
        const int q = k == 0 ? 4 : k - 1 + (k > 4);
        const uint2 r8 = hpel_row8(W.w[0][row + q / 3], mis + hx + q % 3);
        const int sad = warp_sum((int)sad4(r8.y, c8.y, sad4(r8.x, c8.x, 0)));
        if (lane == k) mycost += sad;
    }
    {
        const uint8_t* blk = mb0 + (ptrdiff_t)row * g.ys + hx;
        const uint2 z8 = ld8_unaligned(blk), p8 = ld8_unaligned(blk + (ptrdiff_t)pvy * g.ys + pvx);
        const int s9 = warp_sum((int)sad4(z8.y, c8.y, sad4(z8.x, c8.x, 0)));
        const int s10 = warp_sum((int)sad4(p8.y, c8.y, sad4(p8.x, c8.x, 0)));
        if (lane == 9) mycost += s9;
        if (lane == 10) mycost += s10;
    }
    uint32_t best = warp_min(lane < 11 ? (((uint32_t)mycost << 4) | (uint32_t)lane) : 0xffffffffu);
    const int bvx = __shfl_sync(0xffffffffu, cvx, (int)(best & 15)), bvy = __shfl_sync(0xffffffffu, cvy, (int)(best & 15));
    uint32_t bcost = best >> 4;
    if (bcost < VCP_SUBPEL_SKIP_COST) {   // warp-uniform
        if (lane == 0) {
            b.mv[(size_t)gi * g.nmb + mbi] = make_short2((short)(4 * bvx), (short)(4 * bvy));
            b.mbtype[(size_t)gi * g.nmb + mbi] = VCP_MB_P16;
        }
        return;
    }

    int ox = 0, oy = 0;
    if (!g.hevc || g.hevc_subpel) {   // HEVC: full samples only, or (hevc_subpel) the half-sample step alone
        // half-pel then quarter-pel neighbours from the half-sample planes of the reference, staged
        // once around the best full-pel position
        __syncwarp();
        mis = hpel_window_stage(W, mb0 + (ptrdiff_t)bvy * g.ys + bvx, g.ys, g.ysize, lane, 4);
        const int lane_byte = (row + 1) * 24 + mis + 1 + hx;   // this lane's first sample at displacement 0 inside a plane window
#pragma unroll
        for (int step = 2; step >= 1; step--) {
            if (step == 1 && g.hevc) continue;   // HEVC quarter positions are not averages of half-sample planes
            // lane k (1..8): candidate k -> byte offsets of its two grid samples, and its vector cost
            int o1 = 0, o2 = 0, cq = 0, cr = 0;
            {
                const int k = (lane - 1) & 7, q = k + (k > 3);
                cq = ox + (q % 3 - 1) * step; cr = oy + (q / 3 - 1) * step;
                const HpelPoints h = hpel_points(cq, cr);
                o1 = (hpel_plane(h.x1, h.y1) * 18 + (h.y1 >> 1)) * 24 + (h.x1 >> 1);
                o2 = (hpel_plane(h.x2, h.y2) * 18 + (h.y2 >> 1)) * 24 + (h.x2 >> 1);
            }
            mycost = lam * (vcp_se_len(4 * bvx + cq - pmx) + vcp_se_len(4 * bvy + cr - pmy));
            if (lane == 0) mycost = (int)bcost;
#pragma unroll
            for (int k = 1; k <= 8; k++) {
                const int a1 = lane_byte + __shfl_sync(0xffffffffu, o1, k);
                uint2 p8 = hpel_row8(Ww, a1);
                if (step == 1) {   // quarter positions average two grid samples; half positions are one
                    const int a2 = lane_byte + __shfl_sync(0xffffffffu, o2, k);
                    const uint2 c2 = hpel_row8(Ww, a2);
                    p8 = make_uint2(__vavgu4(p8.x, c2.x), __vavgu4(p8.y, c2.y));
                }
                const int sad = warp_sum((int)sad4(p8.y, c8.y, sad4(p8.x, c8.x, 0)));
                if (lane == k) mycost += sad;
            }
            best = warp_min(lane < 9 ? (((uint32_t)mycost << 4) | (uint32_t)lane) : 0xffffffffu);
            bcost = best >> 4;
            const int bk = (int)(best & 15);
            ox = __shfl_sync(0xffffffffu, bk ? cq : ox, bk);
            oy = __shfl_sync(0xffffffffu, bk ? cr : oy, bk);
        }
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
        const int sh = warp_sum((int)sad4(c8.y, l4, sad4(c8.x, l4, 0u)));
        const int sd = warp_sum((int)sad4(c8.y, d4, sad4(c8.x, d4, 0u)));
        int best_i = sd;
        if (aT && sv < best_i) best_i = sv;
        if (aL && sh < best_i) best_i = sh;
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
void vcp_launch_me_prepass(const VcpGeom& g, const VcpBufs& b, int nframes, int gop, int t, cudaStream_t st, int g0, int g1) {
    if (nframes <= 0) return;
    int nz = t < 0 ? nframes : (nframes - t + gop - 1) / gop;   // GOPs that own a picture t
    if (t >= 0) { if (g1 >= 0 && g1 < nz) nz = g1; nz -= g0; } else g0 = 0;
    if (nz <= 0) return;
    dim3 grid((g.mbw + ME_WARPS - 1) / ME_WARPS, g.mbh, nz);
    me_prepass_kernel<<<grid, ME_WARPS * 32, 0, st>>>(g, b, nframes, gop, t, g0);
}

void vcp_launch_me_refine(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + RF_WARPS - 1) / RF_WARPS, s.ngop);
    me_refine_kernel<<<grid, RF_WARPS * 32, 0, st>>>(g, b, s);
}
