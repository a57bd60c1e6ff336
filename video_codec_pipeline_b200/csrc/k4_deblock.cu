// K4 — in-loop deblocking filter (H.264 8.7), plus the two small passes around it:
//   mbinfo_kernel : motion-vector prediction (8.4.1.3), P_Skip inference (8.4.1.1), mvd;
//                   fully parallel because every final vector is already known.
//   deblock_kernel: the normative filter order is macroblock raster order with vertical edges
//                   before horizontal ones, and each macroblock reads samples already filtered
//                   by its left, top and top-right neighbours.  That is a wavefront with index
//                   d = mx + 2*my; one CTA walks one picture (or one slice when filtering does
//                   not cross slice edges), one warp per macroblock of the front, lanes = the 32
//                   sample rows (16 Y + 8 Cb + 8 Cr) for vertical edges and the 32 sample columns
//                   for horizontal edges, tile staged in shared memory.
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
constexpr int DB_WARPS = 16;

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

__device__ void deblock_mb(const VcpGeom& g, const VcpBufs& b, DbTile& T, int slot, int gi, int mx, int my,
                           bool left_ok, bool top_ok, int qp, int lane) {
    const size_t base = (size_t)gi * g.nmb;
    const int mbi = my * g.mbw + mx;
    uint8_t* Y = b.rec_y + (size_t)slot * g.ysize + g.yoff + (size_t)(16 * my) * g.ys + 16 * mx;
    uint8_t* U = b.rec_u + (size_t)slot * g.csize + g.coff + (size_t)(8 * my) * g.cs + 8 * mx;
    uint8_t* V = b.rec_v + (size_t)slot * g.csize + g.coff + (size_t)(8 * my) * g.cs + 8 * mx;
    // stage tile
    for (int i = lane; i < 100; i += 32) {
        const int r = i / 5, c = i % 5;
        reinterpret_cast<uint32_t*>(&T.Y[r][0])[c] = ld_u32(Y + (ptrdiff_t)(r - 4) * g.ys - 4 + 4 * c);
    }
    for (int i = lane; i < 60; i += 32) {
        const int pl = i / 30, r = (i % 30) / 3, c = i % 3;
        reinterpret_cast<uint32_t*>(&T.C[pl][r][0])[c] = ld_u32((pl ? V : U) + (ptrdiff_t)(r - 2) * g.cs - 4 + 4 * c);
    }
    // boundary strengths: lane -> dir = lane>>4, edge = (lane>>2)&3, segment = lane&3
    {
        const int dir = lane >> 4, ed = (lane >> 2) & 3, k = lane & 3;
        const bool mbedge = ed == 0;
        int bS = 0;
        const bool edge_on = !mbedge || (dir == 0 ? left_ok : top_ok);
        if (edge_on) {
            const size_t po = mbedge ? (dir == 0 ? base + mbi - 1 : base + mbi - g.mbw) : base + mbi;
            const size_t qo = base + mbi;
            const int pt = b.mbtype[po], qt = b.mbtype[qo];
            if (pt == VCP_MB_I16 || qt == VCP_MB_I16) bS = mbedge ? 4 : 3;
            else {
                int pblk, qblk;
                if (dir == 0) { pblk = k * 4 + (mbedge ? 3 : ed - 1); qblk = k * 4 + ed; }
                else { pblk = (mbedge ? 12 : 4 * (ed - 1)) + k; qblk = 4 * ed + k; }
                if (b.nnz[po * 24 + pblk] || b.nnz[qo * 24 + qblk]) bS = 2;
                else if (mbedge) {
                    const short2 pm = b.mv[po], qm = b.mv[qo];
                    bS = (vcp_iabs(pm.x - qm.x) >= 4 || vcp_iabs(pm.y - qm.y) >= 4) ? 1 : 0;
                }
            }
        }
        T.bs[dir][ed][k] = (uint8_t)bS;
    }
    __syncwarp();
    const int qpc = vcp_chroma_qp[qp];
    const int aY = vcp_alpha_tab[qp], bY = vcp_beta_tab[qp], aC = vcp_alpha_tab[qpc], bC = vcp_beta_tab[qpc];
    // vertical edges: lanes 0-15 luma rows, 16-23 Cb rows, 24-31 Cr rows
    if (lane < 16) {
#pragma unroll
        for (int ed = 0; ed < 4; ed++) {
            const int bS = T.bs[0][ed][lane >> 2];
            if (bS) filt_luma(&T.Y[lane + 4][4 + 4 * ed], 1, bS, aY, bY, bS < 4 ? vcp_tc0_tab[qp][bS - 1] : 0);
        }
    } else {
        const int pl = (lane - 16) >> 3, r = lane & 7;
#pragma unroll
        for (int ed = 0; ed < 4; ed += 2) {
            const int bS = T.bs[0][ed][r >> 1];
            if (bS) filt_chroma(&T.C[pl][r + 2][4 + 2 * ed], 1, bS, aC, bC, bS < 4 ? vcp_tc0_tab[qpc][bS - 1] : 0);
        }
    }
    __syncwarp();
    // horizontal edges: lanes 0-15 luma columns, 16-23 Cb columns, 24-31 Cr columns
    if (lane < 16) {
#pragma unroll
        for (int ed = 0; ed < 4; ed++) {
            const int bS = T.bs[1][ed][lane >> 2];
            if (bS) filt_luma(&T.Y[4 + 4 * ed][lane + 4], 24, bS, aY, bY, bS < 4 ? vcp_tc0_tab[qp][bS - 1] : 0);
        }
    } else {
        const int pl = (lane - 16) >> 3, cx = lane & 7;
#pragma unroll
        for (int ed = 0; ed < 4; ed += 2) {
            const int bS = T.bs[1][ed][cx >> 1];
            if (bS) filt_chroma(&T.C[pl][2 + 2 * ed][cx + 4], 12, bS, aC, bC, bS < 4 ? vcp_tc0_tab[qpc][bS - 1] : 0);
        }
    }
    __syncwarp();
    // write back: rows 0..15 all 5 words, rows -3..-1 words 1..4
    for (int i = lane; i < 16 * 5 + 3 * 4; i += 32) {
        int r, c;
        if (i < 80) { r = i / 5; c = i % 5; } else { r = -3 + (i - 80) / 4; c = 1 + (i - 80) % 4; }
        *reinterpret_cast<uint32_t*>(Y + (ptrdiff_t)r * g.ys - 4 + 4 * c) = reinterpret_cast<const uint32_t*>(&T.Y[r + 4][0])[c];
    }
    for (int i = lane; i < 2 * (8 * 3 + 2 * 2); i += 32) {
        const int pl = i / 28, j = i % 28;
        int r, c;
        if (j < 24) { r = j / 3; c = j % 3; } else { r = -2 + (j - 24) / 2; c = 1 + (j - 24) % 2; }
        *reinterpret_cast<uint32_t*>((pl ? V : U) + (ptrdiff_t)r * g.cs - 4 + 4 * c) =
            reinterpret_cast<const uint32_t*>(&T.C[pl][r + 2][0])[c];
    }
    __syncwarp();
}

// grid: x = wavefront domain (1, or slices when deblock_idc == 2), y = GOP
__global__ void __launch_bounds__(DB_WARPS * 32) deblock_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ DbTile tiles[DB_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.y;
    const int n = vcp_frame_of(s, gi);
    const int slot = vcp_rec_slot(s, gi, s.t);
    const int qp = b.qp[n];
    int r0 = 0, r1 = g.mbh;
    if (g.deblock_idc == 2) {
        r0 = vcp_slice_first_row(blockIdx.x, g.slices, g.mbh);
        r1 = (int)blockIdx.x + 1 < g.slices ? vcp_slice_first_row(blockIdx.x + 1, g.slices, g.mbh) : g.mbh;
    }
    const int rows = r1 - r0;
    const int nwave = g.mbw + 2 * (rows - 1);
    for (int d = 0; d < nwave; d++) {
        // macroblocks (mx, r0+k) with mx = d - 2k
        int k0 = (d - (g.mbw - 1) + 1) >> 1; if (k0 < 0) k0 = 0;
        int k1 = d >> 1; if (k1 > rows - 1) k1 = rows - 1;
        for (int k = k0 + warp; k <= k1; k += DB_WARPS) {
            const int mx = d - 2 * k, my = r0 + k;
            deblock_mb(g, b, tiles[warp], slot, gi, mx, my, mx > 0, my > r0, qp, lane);
        }
        __syncthreads();
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
    dim3 grid(g.deblock_idc == 2 ? g.slices : 1, s.ngop);
    deblock_kernel<<<grid, DB_WARPS * 32, 0, st>>>(g, b, s);
}

void vcp_launch_pad(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    const int ny = 2 * VCP_PAD * ((g.cw + 2 * VCP_PAD) / 16) + 2 * g.ch * (VCP_PAD / 16);
    const int nc = 2 * VCP_PADC * ((g.cw / 2 + 2 * VCP_PADC) / 8) + 2 * (g.ch / 2) * (VCP_PADC / 8);
    dim3 grid((ny + 2 * nc + 255) / 256, s.ngop);
    pad_kernel<<<grid, 256, 0, st>>>(g, b, s, ny, nc);
}
