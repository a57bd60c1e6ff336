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
#include <cstdlib>

#include "vcp_dev.cuh"

#define VCP_TAB static __device__ const
#include "h264_tables.h"

namespace {

// ---- mbinfo ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mbinfo_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    const int mbi = blockIdx.x * blockDim.x + threadIdx.x;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    const size_t base = (size_t)gi * g.nmb;
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    const int row0 = vcp_row_first(b, my);
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

struct __align__(16) DbTile {
    uint8_t Y[20][24];     // rows y=-4..15 (idx y+4), cols x=-4..15 (idx x+4)
    uint8_t C[2][10][12];  // rows y=-2..7 (idx y+2), cols x=-4..7 (idx x+4)
    uint8_t bs[2][4][4];   // [dir][edge][segment]
};

// One line across an edge, in registers: v[0..7] = p3 p2 p1 p0 q0 q1 q2 q3.  Luma and chroma lanes run the SAME
// instruction stream (chroma = the luma filter with its p1 / q1 / three-tap branches switched off: 8.7.2.3 and 8.7.2.4 are
// written that way), results are selected, nothing branches on sample values: the two plane types no longer serialise
// inside the warp and the filter conditions cost no branch resolution on the wavefront's critical path.
//   strong: warp-uniform, true when some lane of the warp has bS 4 (intra macroblock edges: rare in P pictures)
__device__ __forceinline__ void filt_edge(int* v, int bS, bool chroma, int alpha, int beta, int tc0, bool strong) {
    const int p3 = v[0], p2 = v[1], p1 = v[2], p0 = v[3], q0 = v[4], q1 = v[5], q2 = v[6], q3 = v[7];
    const int dpq = vcp_iabs(p0 - q0);
    const bool on = bS > 0 && dpq < alpha && vcp_iabs(p1 - p0) < beta && vcp_iabs(q1 - q0) < beta;
    const bool apb = !chroma && vcp_iabs(p2 - p0) < beta, aqb = !chroma && vcp_iabs(q2 - q0) < beta;
    // bS < 4
    const int tc = tc0 + (chroma ? 1 : (int)apb + (int)aqb);
    const int d = vcp_clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
    const int avg = (p0 + q0 + 1) >> 1;
    int n1 = apb ? p1 + vcp_clip3(-tc0, tc0, (p2 + avg - 2 * p1) >> 1) : p1;   // p1'
    int n2 = p2;
    int n3 = vcp_clip255(p0 + d);                                             // p0'
    int n4 = vcp_clip255(q0 - d);                                             // q0'
    int n5 = aqb ? q1 + vcp_clip3(-tc0, tc0, (q2 + avg - 2 * q1) >> 1) : q1;   // q1'
    int n6 = q2;
    if (strong) {
        const bool s4 = bS == 4;
        const bool small = dpq < ((alpha >> 2) + 2);
        const bool ps = apb && small, qs = aqb && small;
        const int sp0 = ps ? (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3 : (2 * p1 + p0 + q1 + 2) >> 2;
        const int sp1 = ps ? (p2 + p1 + p0 + q0 + 2) >> 2 : p1;
        const int sp2 = ps ? (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3 : p2;
        const int sq0 = qs ? (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3 : (2 * q1 + q0 + p1 + 2) >> 2;
        const int sq1 = qs ? (p0 + q0 + q1 + q2 + 2) >> 2 : q1;
        const int sq2 = qs ? (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3 : q2;
        if (s4) { n1 = sp1; n2 = sp2; n3 = sp0; n4 = sq0; n5 = sq1; n6 = sq2; }
    }
    if (on) { v[1] = n2; v[2] = n1; v[3] = n3; v[4] = n4; v[5] = n5; v[6] = n6; }
}

// Everything one macroblock needs that does not depend on the row above, fetched one iteration
// ahead as INDEPENDENT loads (a warp executes in order: a dependent chain would stall it).
struct DbPrefetch {
    uint32_t y0, y1;   // luma: lane -> row = lane>>1, words 2*(lane&1), +1
    uint32_t c;        // chroma: lane -> plane = lane>>4, row = (lane>>1)&7, word lane&1
    uint8_t pt, qt, pn, qn;   // types and nnz of the two 4x4 blocks across this lane's edge segment
    uint8_t qf;               // modes byte of the current macroblock (bit 7: transform_size_8x8_flag)
    short2 pm, qm;
};

// per-lane addresses of the first macroblock of the row; advancing by one macroblock is a
// constant stride, so the loop body carries no address arithmetic
struct DbLaneAddr {
    const uint8_t* y;     // luma: row lane>>1, 8 bytes at 8*(lane&1)
    const uint8_t* c;     // chroma: plane lane>>4, row (lane>>1)&7, 4 bytes at 4*(lane&1)
    const uint8_t* qt;    // mbtype of the current macroblock
    const uint8_t* qn;    // nnz of the q-side block
    const uint8_t* qf;    // modes of the current macroblock
    const short2* qm;     // mv of the current macroblock
    int p_mb_off;         // macroblock offset of the p side (0, -1 or -mbw)
    int pn_off;           // nnz byte offset of the p-side block relative to qn
    bool needs_left;      // the edge exists only when mx > 0
    bool off;             // edge never filtered (top edge at a picture / slice boundary)
};

__device__ __forceinline__ DbLaneAddr db_lane_addr(const VcpGeom& g, const VcpBufs& b, const uint8_t* Y, const uint8_t* U,
                                                   const uint8_t* V, size_t base, int my, bool top_ok, int lane) {
    DbLaneAddr a;
    a.y = Y + (size_t)(lane >> 1) * g.ys + 8 * (lane & 1);
    a.c = ((lane >> 4) ? V : U) + (size_t)((lane >> 1) & 7) * g.cs + 4 * (lane & 1);
    const int dir = lane >> 4, ed = (lane >> 2) & 3, k = lane & 3;
    const bool mbedge = ed == 0;
    int pblk, qblk;
    if (dir == 0) { pblk = k * 4 + (mbedge ? 3 : ed - 1); qblk = k * 4 + ed; }
    else { pblk = (mbedge ? 12 : 4 * (ed - 1)) + k; qblk = 4 * ed + k; }
    const size_t qo = base + (size_t)my * g.mbw;
    a.qt = b.mbtype + qo;
    a.qn = b.nnz + qo * 24 + qblk;
    a.qf = b.modes + qo;
    a.qm = b.mv + qo;
    a.needs_left = mbedge && dir == 0;
    a.off = mbedge && dir == 1 && !top_ok;
    a.p_mb_off = !mbedge || a.off ? 0 : (dir == 0 ? -1 : -g.mbw);
    a.pn_off = a.p_mb_off * 24 + pblk - qblk;
    return a;
}

__device__ __forceinline__ DbPrefetch db_prefetch(const DbLaneAddr& a, int mx) {
    DbPrefetch f;
    const uint2 yy = *reinterpret_cast<const uint2*>(a.y + 16 * mx);
    f.y0 = yy.x; f.y1 = yy.y;
    f.c = ld_u32(a.c + 8 * mx);
    const bool dead = a.off || (a.needs_left && mx == 0);
    const int po = dead ? 0 : a.p_mb_off;
    f.qt = a.qt[mx]; f.pt = a.qt[mx + po];
    f.qn = a.qn[24 * mx]; f.pn = a.qn[24 * mx + (dead ? 0 : a.pn_off)];
    f.qf = a.qf[mx];
    f.qm = a.qm[mx]; f.pm = a.qm[mx + po];
    if (dead) f.pt = 0xff;   // marks "edge not filtered"
    return f;
}

__device__ __forceinline__ int db_strength(const DbPrefetch& f, int lane) {
    const bool mbedge = ((lane >> 2) & 3) == 0;
    if (f.pt == 0xff) return 0;
    // transform_size_8x8_flag: luma edges 1 and 3 are not transform block edges (chroma takes its
    // strengths from edges 0 and 2 only)
    if ((lane & 4) && f.qt == VCP_MB_P16 && (f.qf & 0x80)) return 0;
    if (f.pt == VCP_MB_I16 || f.qt == VCP_MB_I16) return mbedge ? 4 : 3;
    if (f.pn || f.qn) return 2;
    if (mbedge && (vcp_iabs(f.pm.x - f.qm.x) >= 4 || vcp_iabs(f.pm.y - f.qm.y) >= 4)) return 1;
    return 0;
}

// Inter-row hand-off.  __threadfence() is MEMBAR.SC.GPU + CCTL.IVALL on sm_100a (sequentially
// consistent fence plus a full L1 invalidate, by every lane); the row-to-row link only needs
// release/acquire at gpu scope, issued by ONE lane after a warp barrier (bar.warp.sync orders the
// other lanes' stores before it), and the consumer reads the published samples with ld.cg (L2).
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ uint32_t ldcg_u32(const uint8_t* p) { return __ldcg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
    return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}

// One CTA = one band of up to 16 consecutive macroblock rows of one picture (one warp per row).
//   * rows inside a band hand their bottom samples to the row below through a small ring in
//     SHARED memory guarded by cta-scope flags (no gpu-scope fence, ~100 cycles per hand-off);
//   * only the first row of a band reads the previous band's samples from HBM (ld.cg) behind a
//     gpu-scope progress counter, which the last row of a band publishes every DB_PUB macroblocks.
// Bands are handed out by an atomic ticket in top-to-bottom order, so every dependency points at
// a CTA that is already running.
constexpr int DB_RING = 4;      // ring slots per row
constexpr int DB_SLOT = 96;     // 4 luma rows x 16 + 2 planes x 2 chroma rows x 8
constexpr int DB_PUB = 4;       // gpu-scope publish interval of a band's last row

struct DbShared {
    DbTile* tiles;                // [BH]
    uint8_t* ring;                // [BH][DB_RING][DB_SLOT]
    volatile int* done;           // [BH] macroblocks complete (samples final, written, stashed)
    volatile int* taken;          // [BH] ring slots the row has finished reading from the row above
};

// MAXT only steers the register allocation (512 threads are launched at most): 512 -> 126 registers, 768 -> 80, 1024 -> 64.
// A CTA of 16 warps at 126 registers holds the SM's whole register file for the ~0.7 ms a picture band takes, at an issue
// rate of 0.3 per cycle: nothing else can run beside it.  Fewer registers leave room for the other GOP groups' kernels.
template <int MAXT>
__global__ void __launch_bounds__(MAXT) deblock_kernel(VcpGeom g, VcpBufs b, VcpStep s, int BH, int nbands,
                                                        int* __restrict__ ticket, int* __restrict__ progress, unsigned spin_ns) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int my_ticket;
    DbShared sh;
    sh.tiles = reinterpret_cast<DbTile*>(smem_raw);
    sh.ring = smem_raw + (size_t)BH * sizeof(DbTile);
    sh.done = reinterpret_cast<volatile int*>(sh.ring + (size_t)BH * DB_RING * DB_SLOT);
    sh.taken = sh.done + BH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) my_ticket = atomicAdd(ticket, 1);
    if (threadIdx.x < 2 * BH) sh.done[threadIdx.x] = 0;   // done[] and taken[] are contiguous
    __syncthreads();
    const int gi = s.g0 + my_ticket / nbands, band = my_ticket % nbands;
    const int my = band * BH + warp;
    if (my >= g.mbh || my_ticket >= s.ngop * nbands) return;
    const int n = vcp_frame_of(s, gi);
    const int slot = vcp_rec_slot(s, gi, s.t);
    const int qp = b.qp[n];
    const int qpc = vcp_chroma_qp[qp];
    const bool lum = lane < 16;
    const int alpha = lum ? vcp_alpha_tab[qp] : vcp_alpha_tab[qpc], beta = lum ? vcp_beta_tab[qp] : vcp_beta_tab[qpc];
    int tc[3];
#pragma unroll
    for (int i = 0; i < 3; i++) tc[i] = lum ? vcp_tc0_tab[qp][i] : vcp_tc0_tab[qpc][i];
    bool top_ok = my > 0;
    if (g.deblock_idc == 2) top_ok = my > vcp_row_first(b, my);
    const bool top_smem = warp > 0;                                   // row above is in this CTA
    const bool has_below = my + 1 < g.mbh;
    const bool below_smem = has_below && warp + 1 < BH;               // consumer is in this CTA
    const bool below_gmem = has_below && !below_smem;                 // consumer is the next band
    const size_t base = (size_t)gi * g.nmb;
    uint8_t* Y = vcp_rec_luma(b, g, slot) + g.yoff + (size_t)(16 * my) * g.ys;
    uint8_t* U = b.rec_u + (size_t)slot * g.csize + g.coff + (size_t)(8 * my) * g.cs;
    uint8_t* V = b.rec_v + (size_t)slot * g.csize + g.coff + (size_t)(8 * my) * g.cs;
    const int* prog_up = progress + (size_t)gi * g.mbh + my - 1;
    int* prog_me = progress + (size_t)gi * g.mbh + my;
    DbTile& T = sh.tiles[warp];
    uint8_t* ring_me = sh.ring + (size_t)warp * DB_RING * DB_SLOT;          // what I hand down
    const uint8_t* ring_up = sh.ring + (size_t)(warp - 1) * DB_RING * DB_SLOT;  // what I receive
    const DbLaneAddr la = db_lane_addr(g, b, Y, U, V, base, my, top_ok, lane);
    DbPrefetch f = db_prefetch(la, 0);
    int seen_up = 0;
    uint32_t prev_nz = 0u;
    for (int mx = 0; mx < g.mbw; mx++) {
        // (1) own samples and boundary strengths -> tile
        const int mybs = db_strength(f, lane);
        {
            uint32_t* row = reinterpret_cast<uint32_t*>(&T.Y[4 + (lane >> 1)][4 + 8 * (lane & 1)]);
            row[0] = f.y0; row[1] = f.y1;
            *reinterpret_cast<uint32_t*>(&T.C[lane >> 4][2 + ((lane >> 1) & 7)][4 + 4 * (lane & 1)]) = f.c;
            T.bs[lane >> 4][(lane >> 2) & 3][lane & 3] = (uint8_t)mybs;
        }
        const uint32_t nzmask = __ballot_sync(0xffffffffu, mybs != 0);
        const bool top_used = (nzmask & 0x000f0000u) != 0;
        if (mx + 1 < g.mbw) f = db_prefetch(la, mx + 1);
        // (2) the row above must have finished macroblock mx (incl. the vertical edge of mx+1)
        if (top_used) {
            const int need = mx + 1;
            if (seen_up < need) {
                int v = 0;
                if (lane == 0) {
                    // a waiting row yields its issue slots: the SM is shared with other GOP groups' kernels
                    if (top_smem) { while ((v = sh.done[warp - 1]) < need) { if (spin_ns) __nanosleep(spin_ns); } __threadfence_block(); }
                    else { while ((v = ld_relaxed_gpu(prog_up)) < need) __nanosleep(20); fence_acq_rel_gpu(); }
                }
                seen_up = __shfl_sync(0xffffffffu, v, 0);
            }
            // (3) its bottom rows -> tile top rows
            if (top_smem) {
                const uint32_t* sl = reinterpret_cast<const uint32_t*>(ring_up + (mx % DB_RING) * DB_SLOT);
                if (lane < 16) *reinterpret_cast<uint32_t*>(&T.Y[lane >> 2][4 + 4 * (lane & 3)]) = sl[lane];
                else if (lane < 24) {
                    const int pl = (lane - 16) >> 2, r = (lane >> 1) & 1, w = lane & 1;
                    *reinterpret_cast<uint32_t*>(&T.C[pl][r][4 + 4 * w]) = sl[lane];
                }
            } else if (lane < 16) {
                *reinterpret_cast<uint32_t*>(&T.Y[lane >> 2][4 + 4 * (lane & 3)]) =
                    ldcg_u32(Y + (ptrdiff_t)((lane >> 2) - 4) * g.ys + 16 * mx + 4 * (lane & 3));
            } else if (lane < 24) {
                const int pl = (lane - 16) >> 2, r = (lane >> 1) & 1, w = lane & 1;
                *reinterpret_cast<uint32_t*>(&T.C[pl][r][4 + 4 * w]) =
                    ldcg_u32((pl ? V : U) + (ptrdiff_t)(r - 2) * g.cs + 8 * mx + 4 * w);
            }
        }
        __syncwarp();
        if (top_smem && lane == 0) sh.taken[warp] = mx + 1;   // slot mx may be recycled by the row above
        const bool strong_any = __any_sync(0xffffffffu, mybs == 4);
        if (nzmask & 0x0000ffffu) {
            // vertical edges, whole sample row in registers: lanes 0-15 luma rows (20 samples, edges at 4 8 12 16), lanes 16-31
            // chroma rows (12 samples, edges at 4 and 8 = slots 0 and 1); one code path for both
            const int pl = (lane - 16) >> 3, r = lane & 7;
            uint32_t* row = lum ? reinterpret_cast<uint32_t*>(&T.Y[lane + 4][0]) : reinterpret_cast<uint32_t*>(&T.C[pl][r + 2][0]);
            const uint8_t* bsr = lum ? &T.bs[0][0][lane >> 2] : &T.bs[0][0][r >> 1];   // luma: edge e at +4e; chroma: edge 2e at +8e
            const int bstep = lum ? 4 : 8;
            int v[20];
#pragma unroll
            for (int w = 0; w < 5; w++) {
                const uint32_t x = (lum || w < 3) ? row[w] : 0u;
                v[4 * w] = x & 255; v[4 * w + 1] = (x >> 8) & 255; v[4 * w + 2] = (x >> 16) & 255; v[4 * w + 3] = x >> 24;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int bS = (lum || e < 2) ? bsr[bstep * e] : 0;
                // an edge slot no lane filters (inter macroblocks without coefficients on either side) is skipped by the whole warp
                if (__any_sync(0xffffffffu, bS > 0))
                    filt_edge(&v[4 * e], bS, !lum, alpha, beta, bS == 1 ? tc[0] : (bS == 2 ? tc[1] : tc[2]), strong_any);
            }
#pragma unroll
            for (int w = 0; w < 5; w++)
                if (lum || w < 3) row[w] = pack4(v[4 * w], v[4 * w + 1], v[4 * w + 2], v[4 * w + 3]);
        }
        __syncwarp();
        if (nzmask & 0xffff0000u) {
            // horizontal edges, whole sample column in registers: lanes 0-15 luma columns (tile rows 0..19 = y -4..15), lanes
            // 16-31 chroma columns (tile rows 0..9 = y -2..7, held in v[2..11] so that the edges sit where luma's do)
            const int pl = (lane - 16) >> 3, cx = lane & 7;
            uint8_t* col = lum ? &T.Y[0][lane + 4] : &T.C[pl][0][cx + 4];
            const int pitch = lum ? 24 : 12;
            const uint8_t* bsr = lum ? &T.bs[1][0][lane >> 2] : &T.bs[1][0][cx >> 1];
            const int bstep = lum ? 4 : 8;
            int v[20];
#pragma unroll
            for (int r = 0; r < 20; r++) {
                // chroma: v[r] = tile row r - 2 (rows above the tile do not exist: any value, never used by the chroma filter)
                const int tr = lum ? r : (r < 2 ? 0 : (r < 12 ? r - 2 : 9));
                v[r] = (lum || r < 12) ? col[pitch * tr] : 0;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int bS = (lum || e < 2) ? bsr[bstep * e] : 0;
                // an edge slot no lane filters (inter macroblocks without coefficients on either side) is skipped by the whole warp
                if (__any_sync(0xffffffffu, bS > 0))
                    filt_edge(&v[4 * e], bS, !lum, alpha, beta, bS == 1 ? tc[0] : (bS == 2 ? tc[1] : tc[2]), strong_any);
            }
#pragma unroll
            for (int r = 1; r < 19; r++) {
                if (lum) col[24 * r] = (uint8_t)v[r];
                else if (r >= 3 && r <= 10) col[12 * (r - 2)] = (uint8_t)v[r];
            }
        }
        __syncwarp();
        const bool last = mx + 1 == g.mbw;
        // (4) hand-down ring: bottom 4 luma / 2 chroma rows.  Words 0..2 of macroblock mx are final
        //     now; word 3 (x=12..15) of macroblock mx-1 became final with this vertical-edge pass.
        if (below_smem) {
            if (mx >= DB_RING && lane == 0) while (sh.taken[warp + 1] < mx - DB_RING + 1) { if (spin_ns) __nanosleep(spin_ns); }
            __syncwarp();
            uint32_t* cur = reinterpret_cast<uint32_t*>(ring_me + (mx % DB_RING) * DB_SLOT);
            uint32_t* prv = reinterpret_cast<uint32_t*>(ring_me + ((mx + DB_RING - 1) % DB_RING) * DB_SLOT);
            if (lane < 16) {
                const int r = lane >> 2, w = lane & 3;           // luma row 12+r, word w
                const uint32_t* trow = reinterpret_cast<const uint32_t*>(&T.Y[16 + r][0]);
                if (w < 3 || last) cur[lane] = trow[1 + w];
                if (w == 3 && mx > 0) prv[lane] = trow[0];
            } else if (lane < 24) {
                const int pl = (lane - 16) >> 2, r = (lane >> 1) & 1, w = lane & 1;   // chroma row 6+r
                const uint32_t* trow = reinterpret_cast<const uint32_t*>(&T.C[pl][8 + r][0]);
                if (w == 0 || last) cur[lane] = trow[1 + w];
                if (w == 1 && mx > 0) prv[lane] = trow[0];
            }
        }
        // (5) write back.  Luma rows 0..15: x=-4..11 now (x=12..15 wait for the next vertical
        //     edge; on the last macroblock they go out too); top rows -3..-1: x=0..15.
        //     Nothing to store when neither this macroblock nor the previous one (whose last four columns go out
        //     now) filtered anything: the tile still holds what the picture holds.
        const bool wb = (nzmask | prev_nz) != 0u;
        prev_nz = nzmask;
        if (wb) {
            const int r = lane >> 1, h = lane & 1;
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
        if (wb) {
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
        // (6) publish: after this iteration's stores, macroblocks 0..mx-1 (all, on the last one) are
        //     complete.  In-CTA consumers need only cta-scope ordering (their later stores to the
        //     shared bottom rows must land after ours); the next band needs gpu scope.
        const int complete = last ? g.mbw : mx;
        if (below_smem && lane == 0) { __threadfence_block(); sh.done[warp] = complete; }
        if (below_gmem && lane == 0 && (last || (mx % DB_PUB) == DB_PUB - 1)) {
            fence_acq_rel_gpu();
            st_relaxed_gpu(prog_me, complete);
        }
        // (7) the right-most columns become the next macroblock's left neighbour
        if (lum) *reinterpret_cast<uint32_t*>(&T.Y[lane + 4][0]) = *reinterpret_cast<const uint32_t*>(&T.Y[lane + 4][16]);
        else *reinterpret_cast<uint32_t*>(&T.C[(lane - 16) >> 3][2 + (lane & 7)][0]) =
                 *reinterpret_cast<const uint32_t*>(&T.C[(lane - 16) >> 3][2 + (lane & 7)][8]);
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
    const int gi = blockIdx.y + s.g0;
    const int slot = vcp_rec_slot(s, gi, s.t);
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < ny) { pad_plane<16>(vcp_rec_luma(b, g, slot) + g.yoff, g.ys, g.cw, g.ch, VCP_PAD, idx); return; }
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
    // bands of up to 16 macroblock rows, one CTA each (512 threads: no register cap pressure)
    const int nbands = (g.mbh + 15) / 16, BH = (g.mbh + nbands - 1) / nbands;
    const size_t smem = (size_t)BH * (sizeof(DbTile) + DB_RING * DB_SLOT + 2 * sizeof(int));
    // sync words: [0] ticket, [1..] per-row progress (each GOP group has its own region: groups
    // run concurrently on different streams)
    int* sync = b.db_sync + (size_t)s.g0 * (g.mbh + 1);
    cudaMemsetAsync(sync, 0, (size_t)(s.ngop * g.mbh + 1) * sizeof(int), st);
    // function attributes are per device: a process may drive several GPUs from different threads
    static const unsigned spin_ns = [] { const char* e = getenv("VCPENC_DB_SPIN_NS"); return e ? (unsigned)atoi(e) : 0u; }();
    // Register build by how crowded the machine is.  While every band CTA of the batch (all GOP groups' launches together)
    // can have an SM of its own, the 126-register build is faster (1080p step, 80 vs 126 registers: 10 GOPs 68.4 / 57.0 ms,
    // 16: 80.3 / 70.0, 24: 98.5 / 91.3); beyond that the CTAs hold registers the throughput kernels of the other groups
    // need, and 80 wins (32 GOPs: 118.3 / 120.2; 64 registers spill: 128.2).
    static const int forced = [] { const char* e = getenv("VCPENC_DB_REGS"); return e ? atoi(e) : 0; }();
    static const int nsm = [] { int dev = 0, n = 148; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n > 0 ? n : 148; }();
    const int total_gops = (s.nframes + s.gop - 1) / s.gop;
    const int regs = forced ? forced : (total_gops * nbands <= nsm ? 126 : 80);
    int* progress = sync + 1 - (size_t)s.g0 * g.mbh;
    if (regs >= 120) deblock_kernel<512><<<s.ngop * nbands, BH * 32, smem, st>>>(g, b, s, BH, nbands, sync, progress, spin_ns);
    else if (regs >= 72) deblock_kernel<768><<<s.ngop * nbands, BH * 32, smem, st>>>(g, b, s, BH, nbands, sync, progress, spin_ns);
    else deblock_kernel<1024><<<s.ngop * nbands, BH * 32, smem, st>>>(g, b, s, BH, nbands, sync, progress, spin_ns);
}

void vcp_launch_pad(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    const int ny = 2 * VCP_PAD * ((g.cw + 2 * VCP_PAD) / 16) + 2 * g.ch * (VCP_PAD / 16);
    const int nc = 2 * VCP_PADC * ((g.cw / 2 + 2 * VCP_PADC) / 8) + 2 * (g.ch / 2) * (VCP_PADC / 8);
    dim3 grid((ny + 2 * nc + 255) / 256, s.ngop);
    pad_kernel<<<grid, 256, 0, st>>>(g, b, s, ny, nc);
}
