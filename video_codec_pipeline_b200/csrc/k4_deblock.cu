// K4 — in-loop deblocking filter (H.264 8.7), plus the two small passes around it:
//   mbinfo_kernel : motion-vector prediction (8.4.1.3), P_Skip inference (8.4.1.1), mvd;
//                   fully parallel because every final vector is already known.
//   deblock_kernel: the normative filter order is macroblock raster order with vertical edges
//                   before horizontal ones, and each macroblock reads samples already filtered
//                   by its left, top and top-right neighbours.  That is a wavefront with index
//                   d = mx + 2*my.  One warp streams along one macroblock row (left neighbour
//                   stays in shared memory), rows synchronise through progress counters; lanes =
//                   the 32 sample rows (16 Y + 8 Cb + 8 Cr) for vertical edges and the 32 sample
//                   columns for horizontal edges.
//   pad_kernel    : replicates the picture edge into the border so the next frame's motion
//                   vectors may leave the picture.
//
// Replaces x264's deblock inside the ffmpeg child (/root/reference/cmd/consumer.go:376-382);
// bit-identical to oracle/h264_oracle.c (deblock_frame, mvp16, mv_pskip).
#include "vcp_dev.cuh"

#define VCP_TAB static __device__ const
#include "h264_tables.h"

namespace {

// ---- mbinfo ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mbinfo_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    const int mbi = blockIdx.x * blockDim.x + threadIdx.x;
    const int gi = blockIdx.y;
    if (mbi >= g.nmb) return;
    const size_t base = (size_t)gi * g.nmb;
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    const int row0 = vcp_slice_first_row(vcp_slice_of_row(my, g.slices, g.mbh), g.slices, g.mbh);
    // 0:A left 1:B top 2:C top-right 3:D top-left
    const int nx[4] = {mx - 1, mx, mx + 1, mx - 1}, ny[4] = {my, my - 1, my - 1, my - 1};
    bool av[4]; int rf[4], vx[4], vy[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        av[k] = nx[k] >= 0 && nx[k] < g.mbw && ny[k] >= row0;
        rf[k] = -1; vx[k] = 0; vy[k] = 0;
        if (av[k]) {
            const size_t o = base + ny[k] * g.mbw + nx[k];
            if (b.mbtype[o] != VCP_MB_I16) { const short2 v = b.mv[o]; rf[k] = 0; vx[k] = v.x; vy[k] = v.y; }
        }
    }
    const int c = av[2] ? 2 : 3;
    int px, py;
    if (!av[1] && !av[c] && av[0]) { px = vx[0]; py = vy[0]; }
    else {
        const int cnt = (rf[0] == 0) + (rf[1] == 0) + (rf[c] == 0);
        if (cnt == 1) { const int k = rf[0] == 0 ? 0 : (rf[1] == 0 ? 1 : c); px = vx[k]; py = vy[k]; }
        else { px = vcp_median3(vx[0], vx[1], vx[c]); py = vcp_median3(vy[0], vy[1], vy[c]); }
    }
    int sx = px, sy = py;
    if (!av[0] || !av[1] || (rf[0] == 0 && !vx[0] && !vy[0]) || (rf[1] == 0 && !vx[1] && !vy[1])) { sx = 0; sy = 0; }
    const short2 mv = b.mv[base + mbi];
    if (b.mbtype[base + mbi] == VCP_MB_P16 && b.cbp[base + mbi] == 0 && mv.x == sx && mv.y == sy)
        b.mbtype[base + mbi] = VCP_MB_PSKIP;
    b.mvd[base + mbi] = make_short2((short)(mv.x - px), (short)(mv.y - py));
}

// ---- deblocking --------------------------------------------------------------------------------
// Row-streaming wavefront.  One warp owns one macroblock row and walks it left to right; the
// left neighbour is the warp's own previous macroblock (kept in shared memory, no round trip),
// the top neighbour comes from the warp of the row above through global memory, guarded by a
// per-row progress counter: row r may filter macroblock x once row r-1 has finished x+1
// (the top-right macroblock's vertical edge touches the samples our top edge reads).
// Rows are handed out by an atomic ticket in top-to-bottom order, so every dependency points
// at a warp that is already running (no deadlock however the hardware orders CTAs).
constexpr int DB_WARPS = 4;

struct __align__(16) DbTile {
    uint8_t Y[20][24];     // rows y=-4..15 (idx y+4), cols x=-4..15 (idx x+4)
    uint8_t C[2][10][12];  // rows y=-2..7 (idx y+2), cols x=-4..7 (idx x+4)
    uint8_t bs[2][4][4];   // [dir][edge][segment]
};

__device__ __forceinline__ void filt_luma(uint8_t* pix, int xs, int bS, int alpha, int beta, int tc0) {
    const int p0 = pix[-xs], p1 = pix[-2 * xs], p2 = pix[-3 * xs], q0 = pix[0], q1 = pix[xs], q2 = pix[2 * xs];
    if (vcp_iabs(p0 - q0) >= alpha || vcp_iabs(p1 - p0) >= beta || vcp_iabs(q1 - q0) >= beta) return;
    const int ap = vcp_iabs(p2 - p0), aq = vcp_iabs(q2 - q0);
    if (bS < 4) {
        const int tc = tc0 + (ap < beta) + (aq < beta);
        const int d = vcp_clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
        pix[-xs] = (uint8_t)vcp_clip255(p0 + d);
        pix[0] = (uint8_t)vcp_clip255(q0 - d);
        if (ap < beta) pix[-2 * xs] = (uint8_t)(p1 + vcp_clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - 2 * p1) >> 1));
        if (aq < beta) pix[xs] = (uint8_t)(q1 + vcp_clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - 2 * q1) >> 1));
    } else {
        const int p3 = pix[-4 * xs], q3 = pix[3 * xs];
        const bool small = vcp_iabs(p0 - q0) < ((alpha >> 2) + 2);
        if (ap < beta && small) {
            pix[-xs] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            pix[-2 * xs] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            pix[-3 * xs] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else pix[-xs] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        if (aq < beta && small) {
            pix[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            pix[xs] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            pix[2 * xs] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}
__device__ __forceinline__ void filt_chroma(uint8_t* pix, int xs, int bS, int alpha, int beta, int tc0) {
    const int p0 = pix[-xs], p1 = pix[-2 * xs], q0 = pix[0], q1 = pix[xs];
    if (vcp_iabs(p0 - q0) >= alpha || vcp_iabs(p1 - p0) >= beta || vcp_iabs(q1 - q0) >= beta) return;
    if (bS < 4) {
        const int tc = tc0 + 1;
        const int d = vcp_clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
        pix[-xs] = (uint8_t)vcp_clip255(p0 + d);
        pix[0] = (uint8_t)vcp_clip255(q0 - d);
    } else {
        pix[-xs] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

// per-macroblock data fetched one iteration ahead (independent of the row above)
struct DbPrefetch {
    uint32_t y0, y1;   // luma: lane -> row = lane>>1, words 2*(lane&1), +1
    uint32_t c;        // chroma: lane -> plane = lane>>4, row = (lane>>1)&7, word lane&1
    int bs;            // boundary strength of (dir = lane>>4, edge = (lane>>2)&3, seg = lane&3)
};

__device__ __forceinline__ int db_strength(const VcpGeom& g, const VcpBufs& b, size_t base, int mbi, bool left_ok,
                                           bool top_ok, int lane) {
    const int dir = lane >> 4, ed = (lane >> 2) & 3, k = lane & 3;
    const bool mbedge = ed == 0;
    if (mbedge && !(dir == 0 ? left_ok : top_ok)) return 0;
    const size_t po = mbedge ? (dir == 0 ? base + mbi - 1 : base + mbi - g.mbw) : base + mbi;
    const size_t qo = base + mbi;
    const int pt = b.mbtype[po], qt = b.mbtype[qo];
    if (pt == VCP_MB_I16 || qt == VCP_MB_I16) return mbedge ? 4 : 3;
    int pblk, qblk;
    if (dir == 0) { pblk = k * 4 + (mbedge ? 3 : ed - 1); qblk = k * 4 + ed; }
    else { pblk = (mbedge ? 12 : 4 * (ed - 1)) + k; qblk = 4 * ed + k; }
    if (b.nnz[po * 24 + pblk] || b.nnz[qo * 24 + qblk]) return 2;
    if (mbedge) {
        const short2 pm = b.mv[po], qm = b.mv[qo];
        return (vcp_iabs(pm.x - qm.x) >= 4 || vcp_iabs(pm.y - qm.y) >= 4) ? 1 : 0;
    }
    return 0;
}

__device__ __forceinline__ DbPrefetch db_prefetch(const VcpGeom& g, const VcpBufs& b, const uint8_t* Y, const uint8_t* U,
                                                  const uint8_t* V, size_t base, int mx, int my, bool top_ok, int lane) {
    DbPrefetch f;
    const uint2 yy = *reinterpret_cast<const uint2*>(Y + (size_t)(lane >> 1) * g.ys + 16 * mx + 8 * (lane & 1));
    f.y0 = yy.x; f.y1 = yy.y;
    f.c = ld_u32(((lane >> 4) ? V : U) + (size_t)((lane >> 1) & 7) * g.cs + 8 * mx + 4 * (lane & 1));
    f.bs = db_strength(g, b, base, my * g.mbw + mx, mx > 0, top_ok, lane);
    return f;
}

__device__ __forceinline__ uint32_t ldcg_u32(const uint8_t* p) { return __ldcg(reinterpret_cast<const uint32_t*>(p)); }

// grid: x = row groups (claimed by ticket); progress[gop][row] counts finished macroblocks
__global__ void __launch_bounds__(DB_WARPS * 32) deblock_kernel(VcpGeom g, VcpBufs b, VcpStep s, int* __restrict__ ticket,
                                                                 int* __restrict__ progress) {
    __shared__ DbTile tiles[DB_WARPS];
    __shared__ int my_ticket;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) my_ticket = atomicAdd(ticket, 1);
    __syncthreads();
    const int R = my_ticket * DB_WARPS + warp;          // global row index, frame-major
    if (R >= s.ngop * g.mbh) return;
    const int gi = R / g.mbh, my = R % g.mbh;
    const int n = vcp_frame_of(s, gi);
    const int slot = vcp_rec_slot(s, gi, s.t);
    const int qp = b.qp[n];
    const int qpc = vcp_chroma_qp[qp];
    const int aY = vcp_alpha_tab[qp], bY = vcp_beta_tab[qp], aC = vcp_alpha_tab[qpc], bC = vcp_beta_tab[qpc];
    int tcY[3], tcC[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { tcY[i] = vcp_tc0_tab[qp][i]; tcC[i] = vcp_tc0_tab[qpc][i]; }
    bool top_ok = my > 0;
    if (g.deblock_idc == 2) top_ok = my > vcp_slice_first_row(vcp_slice_of_row(my, g.slices, g.mbh), g.slices, g.mbh);
    const size_t base = (size_t)gi * g.nmb;
    uint8_t* Y = b.rec_y + (size_t)slot * g.ysize + g.yoff + (size_t)(16 * my) * g.ys;
    uint8_t* U = b.rec_u + (size_t)slot * g.csize + g.coff + (size_t)(8 * my) * g.cs;
    uint8_t* V = b.rec_v + (size_t)slot * g.csize + g.coff + (size_t)(8 * my) * g.cs;
    volatile int* prog_up = progress + (size_t)gi * g.mbh + my - 1;
    volatile int* prog_me = progress + (size_t)gi * g.mbh + my;
    DbTile& T = tiles[warp];
    DbPrefetch f = db_prefetch(g, b, Y, U, V, base, 0, my, top_ok, lane);
    for (int mx = 0; mx < g.mbw; mx++) {
        // (1) own samples -> tile
        {
            uint32_t* row = reinterpret_cast<uint32_t*>(&T.Y[4 + (lane >> 1)][4 + 8 * (lane & 1)]);
            row[0] = f.y0; row[1] = f.y1;
            *reinterpret_cast<uint32_t*>(&T.C[lane >> 4][2 + ((lane >> 1) & 7)][4 + 4 * (lane & 1)]) = f.c;
            T.bs[lane >> 4][(lane >> 2) & 3][lane & 3] = (uint8_t)f.bs;
        }
        const bool any = __any_sync(0xffffffffu, f.bs != 0);
        const bool top_used = top_ok && __any_sync(0xffffffffu, (lane >> 2) == 4 && f.bs != 0);
        // prefetch the next macroblock while we wait / filter
        if (mx + 1 < g.mbw) f = db_prefetch(g, b, Y, U, V, base, mx + 1, my, top_ok, lane);
        // (2) wait for the row above, (3) fetch its bottom rows
        if (top_used) {
            const int need = mx + 2 < g.mbw + 1 ? mx + 2 : g.mbw + 1;
            if (lane == 0) while (*prog_up < need) __nanosleep(40);
            __syncwarp();
            __threadfence();
            {
                if (lane < 16) {
                    *reinterpret_cast<uint32_t*>(&T.Y[lane >> 2][4 + 4 * (lane & 3)]) =
                        ldcg_u32(Y + (ptrdiff_t)((lane >> 2) - 4) * g.ys + 16 * mx + 4 * (lane & 3));
                } else if (lane < 24) {
                    const int pl = (lane - 16) >> 2, r = (lane >> 1) & 1, w = lane & 1;
                    *reinterpret_cast<uint32_t*>(&T.C[pl][r][4 + 4 * w]) =
                        ldcg_u32((pl ? V : U) + (ptrdiff_t)(r - 2) * g.cs + 8 * mx + 4 * w);
                }
            }
        }
        __syncwarp();
        if (any) {
            // vertical edges: lanes 0-15 luma rows, 16-23 Cb rows, 24-31 Cr rows
            if (lane < 16) {
#pragma unroll
                for (int ed = 0; ed < 4; ed++) {
                    const int bS = T.bs[0][ed][lane >> 2];
                    if (bS) filt_luma(&T.Y[lane + 4][4 + 4 * ed], 1, bS, aY, bY, bS < 4 ? tcY[bS - 1] : 0);
                }
            } else {
                const int pl = (lane - 16) >> 3, r = lane & 7;
#pragma unroll
                for (int ed = 0; ed < 4; ed += 2) {
                    const int bS = T.bs[0][ed][r >> 1];
                    if (bS) filt_chroma(&T.C[pl][r + 2][4 + 2 * ed], 1, bS, aC, bC, bS < 4 ? tcC[bS - 1] : 0);
                }
            }
            __syncwarp();
            // horizontal edges: lanes 0-15 luma columns, 16-23 Cb columns, 24-31 Cr columns
            if (lane < 16) {
#pragma unroll
                for (int ed = 0; ed < 4; ed++) {
                    const int bS = T.bs[1][ed][lane >> 2];
                    if (bS) filt_luma(&T.Y[4 + 4 * ed][lane + 4], 24, bS, aY, bY, bS < 4 ? tcY[bS - 1] : 0);
                }
            } else {
                const int pl = (lane - 16) >> 3, cx = lane & 7;
#pragma unroll
                for (int ed = 0; ed < 4; ed += 2) {
                    const int bS = T.bs[1][ed][cx >> 1];
                    if (bS) filt_chroma(&T.C[pl][2 + 2 * ed][cx + 4], 12, bS, aC, bC, bS < 4 ? tcC[bS - 1] : 0);
                }
            }
            __syncwarp();
        }
        // (6) write back.  Luma rows 0..15: x=-4..11 now (x=12..15 wait for the next vertical
        //     edge; on the last macroblock they go out too); top rows -3..-1: x=0..15.
        const bool last = mx + 1 == g.mbw;
        {
            const int r = lane >> 1, h = lane & 1;   // two words per lane + the pending ones
            uint32_t* dst = reinterpret_cast<uint32_t*>(Y + (size_t)r * g.ys + 16 * mx - 4);
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&T.Y[r + 4][0]);
            if (h == 0) { if (mx > 0) dst[0] = src[0]; dst[1] = src[1]; }
            else { dst[2] = src[2]; dst[3] = src[3]; if (last) dst[4] = src[4]; }
        }
        if (top_used && lane < 12) {
            const int r = lane >> 2, w = lane & 3;
            *reinterpret_cast<uint32_t*>(Y + (ptrdiff_t)(r - 3) * g.ys + 16 * mx + 4 * w) =
                *reinterpret_cast<const uint32_t*>(&T.Y[r + 1][4 + 4 * w]);
        }
        {
            // chroma rows 0..7: x=-4..3 now, x=4..7 next time / on the last macroblock
            const int pl = lane >> 4, r = (lane >> 1) & 7, h = lane & 1;
            uint32_t* dst = reinterpret_cast<uint32_t*>((pl ? V : U) + (size_t)r * g.cs + 8 * mx - 4);
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&T.C[pl][r + 2][0]);
            if (h == 0) { if (mx > 0) dst[0] = src[0]; }
            else { dst[1] = src[1]; if (last) dst[2] = src[2]; }
        }
        if (top_used && lane >= 16 && lane < 24) {
            const int pl = (lane - 16) >> 2, r = (lane >> 1) & 1, w = lane & 1;
            *reinterpret_cast<uint32_t*>((pl ? V : U) + (ptrdiff_t)(r - 2) * g.cs + 8 * mx + 4 * w) =
                *reinterpret_cast<const uint32_t*>(&T.C[pl][r][4 + 4 * w]);
        }
        __syncwarp();
        // (7) the right-most columns become the next macroblock's left neighbour
        if (lane < 16) *reinterpret_cast<uint32_t*>(&T.Y[lane + 4][0]) = *reinterpret_cast<const uint32_t*>(&T.Y[lane + 4][16]);
        else *reinterpret_cast<uint32_t*>(&T.C[(lane - 16) >> 3][2 + (lane & 7)][0]) =
                 *reinterpret_cast<const uint32_t*>(&T.C[(lane - 16) >> 3][2 + (lane & 7)][8]);
        // (8) publish progress
        __threadfence();
        __syncwarp();
        if (lane == 0) *prog_me = last ? g.mbw + 1 : mx + 1;
    }
}

// ---- border extension -----------------------------------------------------------------------------
template <int CH>  // chunk bytes: 16 for luma, 8 for chroma
__device__ __forceinline__ void pad_plane(uint8_t* p, int stride, int w, int h, int pad, int idx) {
    // chunk enumeration: top band, bottom band, left band, right band
    const int rowc = (w + 2 * pad) / CH, band = pad * rowc, side = pad / CH;
    int x0, y;
    if (idx < 2 * band) {
        const int j = idx % band;
        y = idx < band ? -pad + j / rowc : h + j / rowc;
        x0 = -pad + (j % rowc) * CH;
    } else {
        const int j = idx - 2 * band;
        if (j >= 2 * h * side) return;
        y = (j / side) % h;
        x0 = j < h * side ? -pad + (j % side) * CH : w + (j % side) * CH;
    }
    const int yy = y < 0 ? 0 : (y >= h ? h - 1 : y);
    const uint8_t* srow = p + (ptrdiff_t)yy * stride;
    uint8_t* d = p + (ptrdiff_t)y * stride + x0;
    if (x0 >= 0 && x0 + CH <= w) {
        if (CH == 16) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(srow + x0);
        else *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(srow + x0);
    } else {
        const uint32_t v = (uint32_t)srow[x0 < 0 ? 0 : w - 1] * 0x01010101u;
        if (CH == 16) *reinterpret_cast<uint4*>(d) = make_uint4(v, v, v, v);
        else *reinterpret_cast<uint2*>(d) = make_uint2(v, v);
    }
}

__global__ void __launch_bounds__(256) pad_kernel(VcpGeom g, VcpBufs b, VcpStep s, int ny, int nc) {
    const int gi = blockIdx.y;
    const int slot = vcp_rec_slot(s, gi, s.t);
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < ny) { pad_plane<16>(b.rec_y + (size_t)slot * g.ysize + g.yoff, g.ys, g.cw, g.ch, VCP_PAD, idx); return; }
    idx -= ny;
    if (idx < nc) { pad_plane<8>(b.rec_u + (size_t)slot * g.csize + g.coff, g.cs, g.cw / 2, g.ch / 2, VCP_PADC, idx); return; }
    idx -= nc;
    if (idx < nc) pad_plane<8>(b.rec_v + (size_t)slot * g.csize + g.coff, g.cs, g.cw / 2, g.ch / 2, VCP_PADC, idx);
}

}  // namespace

void vcp_launch_mbinfo(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + 127) / 128, s.ngop);
    mbinfo_kernel<<<grid, 128, 0, st>>>(g, b, s);
}

void vcp_launch_deblock(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    if (g.deblock_idc == 1) return;
    // sync words: [0] ticket, [1..] per-row progress
    const int rows = s.ngop * g.mbh;
    cudaMemsetAsync(b.db_sync, 0, (size_t)(rows + 1) * sizeof(int), st);
    deblock_kernel<<<(rows + DB_WARPS - 1) / DB_WARPS, DB_WARPS * 32, 0, st>>>(g, b, s, b.db_sync, b.db_sync + 1);
}

void vcp_launch_pad(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    const int ny = 2 * VCP_PAD * ((g.cw + 2 * VCP_PAD) / 16) + 2 * g.ch * (VCP_PAD / 16);
    const int nc = 2 * VCP_PADC * ((g.cw / 2 + 2 * VCP_PADC) / 8) + 2 * (g.ch / 2) * (VCP_PADC / 8);
    dim3 grid((ny + 2 * nc + 255) / 256, s.ngop);
    pad_kernel<<<grid, 256, 0, st>>>(g, b, s, ny, nc);
}
