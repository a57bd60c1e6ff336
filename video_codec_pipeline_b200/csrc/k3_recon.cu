// K3 — prediction, 4x4 integer transform, quantisation, dequantisation, inverse transform and
// reconstruction (H.264 8.3, 8.4.2, 8.5).  One warp per macroblock; lanes 0-15 own the luma
// 4x4 blocks (coding order), lanes 16-19 the Cb and 20-23 the Cr blocks.
//
//   p_recon_kernel : P_L0_16x16, fully parallel over the macroblocks of all resident GOPs.
//   i_recon_kernel : Intra16x16, wavefront (anti-diagonal) inside one CTA per (frame, slice),
//                    because intra prediction reads the reconstructed left/top neighbours.
//
// Replaces x264's dct/quant/predict inside the ffmpeg child
// (/root/reference/cmd/consumer.go:376-382).  Bit-identical to oracle/h264_oracle.c
// (encode_p_mb, encode_i_mb, encode_chroma); the FFmpeg decoder must reproduce `rec_*`.
#include "vcp_dev.cuh"
#include "vcp_luma_interp.cuh"
#include "vcp_transform.cuh"

namespace {

// coding-order luma block index -> position
__device__ __forceinline__ int blk_x4(int b) { return (b & 1) | ((b >> 1) & 2); }
__device__ __forceinline__ int blk_y4(int b) { return ((b >> 1) & 1) | ((b >> 2) & 2); }
__device__ __forceinline__ int blk_ras(int b) { return blk_y4(b) * 4 + blk_x4(b); }

// ---- shared tail: chroma residual of lanes 16..23, writes levels / nnz / recon ---------------
// pred4[r] = the lane's four predicted rows (packed 4 px); returns per-lane nz of the AC part
constexpr int T8_STRIDE = 72;   // ints per 8x8 block: 64 + 8 of padding keeps the four blocks on different banks
struct __align__(16) McScratch {
    uint8_t pred[16][16];
    int t8[4][T8_STRIDE];   // 8x8 transform staging (High profile): one 8x8 block per eight lanes
    int16_t lv8[256];       // levels of the four 8x8 blocks in record order, written out as 16-byte vectors
};
// per-CTA copies of the 8x8 tables: indexed per lane, so shared memory rather than constant
struct T8Tables {
    uint8_t cls[64], izz[64];
    uint16_t mf[6][6];
    uint8_t v[6][6];
};

// ---- 8x8 transform path of an inter macroblock (High profile) --------------------------------------
__device__ __forceinline__ void fdct8_1d(const int x[8], int y[8]) {
    const int a0 = x[0] + x[7], a1 = x[1] + x[6], a2 = x[2] + x[5], a3 = x[3] + x[4];
    const int b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
    const int a4 = x[0] - x[7], a5 = x[1] - x[6], a6 = x[2] - x[5], a7 = x[3] - x[4];
    const int b4 = a5 + a6 + ((a4 >> 1) + a4), b5 = a4 - a7 - ((a6 >> 1) + a6);
    const int b6 = a4 + a7 - ((a5 >> 1) + a5), b7 = a5 - a6 + ((a7 >> 1) + a7);
    y[0] = b0 + b1; y[1] = b4 + (b7 >> 2); y[2] = b2 + (b3 >> 1); y[3] = b5 + (b6 >> 2);
    y[4] = b0 - b1; y[5] = b6 - (b5 >> 2); y[6] = (b2 >> 1) - b3; y[7] = (b4 >> 2) - b7;
}
__device__ __forceinline__ void idct8_1d(const int d[8], int o[8]) {   // 8.5.13
    const int a0 = d[0] + d[4], a4 = d[0] - d[4], a2 = (d[2] >> 1) - d[6], a6 = d[2] + (d[6] >> 1);
    const int b0 = a0 + a6, b2 = a4 + a2, b4 = a4 - a2, b6 = a0 - a6;
    const int a1 = -d[3] + d[5] - d[7] - (d[7] >> 1), a3 = d[1] + d[7] - d[3] - (d[3] >> 1);
    const int a5 = -d[1] + d[7] + d[5] + (d[5] >> 1), a7 = d[3] + d[5] + d[1] + (d[1] >> 1);
    const int b1 = a1 + (a7 >> 2), b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5, b7 = a7 - (a1 >> 2);
    o[0] = b0 + b7; o[1] = b2 + b5; o[2] = b4 + b3; o[3] = b6 + b1;
    o[4] = b6 - b1; o[5] = b4 - b3; o[6] = b2 - b5; o[7] = b0 - b7;
}

// 4x4 or 8x8?  Sums of absolute Hadamard coefficients of the residual (oracle: prefer_8x8).  Lanes 0..15
// hold the residual of their 4x4 block; the 8x8 Hadamard of [[A,B],[C,D]] is the 4x4 Hadamard of
// A+-B+-C+-D, i.e. two shuffle butterflies among the four lanes of an 8x8 block.
__device__ __forceinline__ bool prefer_8x8(const uint8_t* __restrict__ src, int stride, const uint32_t predw[4], int lane) {
    int h[16];
    int c4 = 0, c8 = 0;
    {
        int d[16], t[16];
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const uint32_t s4 = lane < 16 ? ld_u32(src + (size_t)y * stride) : 0u, p4 = lane < 16 ? predw[y] : 0u;
#pragma unroll
            for (int x = 0; x < 4; x++) d[4 * y + x] = (int)((s4 >> (8 * x)) & 255) - (int)((p4 >> (8 * x)) & 255);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int a = d[4 * i] + d[4 * i + 1], bb = d[4 * i] - d[4 * i + 1], c = d[4 * i + 2] + d[4 * i + 3], e = d[4 * i + 2] - d[4 * i + 3];
            t[4 * i] = a + c; t[4 * i + 1] = bb + e; t[4 * i + 2] = a - c; t[4 * i + 3] = bb - e;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int a = t[i] + t[4 + i], bb = t[i] - t[4 + i], c = t[8 + i] + t[12 + i], e = t[8 + i] - t[12 + i];
            h[i] = a + c; h[4 + i] = bb + e; h[8 + i] = a - c; h[12 + i] = bb - e;
        }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        c4 += vcp_iabs(h[i]);
        const int p1 = __shfl_xor_sync(0xffffffffu, h[i], 1);
        const int u = (lane & 1) ? p1 - h[i] : h[i] + p1;          // A+B | A-B | C+D | C-D
        const int p2 = __shfl_xor_sync(0xffffffffu, u, 2);
        const int v = (lane & 2) ? p2 - u : u + p2;                // the four sign combinations
        c8 += vcp_iabs(v);
    }
    const int cost4 = warp_sum(lane < 16 ? c4 : 0), cost8 = warp_sum(lane < 16 ? c8 : 0);
    return vcp_prefer_8x8(cost4, cost8) != 0;
}

// Luma of an inter macroblock as four 8x8 blocks: eight lanes per block (k = lane >> 3, q = lane & 7); a lane
// owns row q in the row passes and column q in the column passes; the passes meet in shared memory, each row
// rotated by its index so that row and column accesses are both conflict-free.  The inverse runs rows first,
// then columns, as 8.5.13 prescribes.  S.pred holds the prediction on entry and the reconstruction on exit.
// Returns the luma cbp.
__device__ __forceinline__ uint32_t luma8x8_transform(const VcpGeom& g, const VcpBufs& b, McScratch& S, const T8Tables& TT,
                                                      int n, int gi, int mbi, int mx, int my, int qp, int lane) {
    const int k = lane >> 3, q = lane & 7;
    const int bx = (k & 1) * 8, by = (k >> 1) * 8;
    int* T = S.t8[k];
    const uint8_t* src = b.src_y + (size_t)n * g.ysize + g.yoff + (size_t)(16 * my + by) * g.ys + 16 * mx + bx;
    {
        const uint2 s8 = *reinterpret_cast<const uint2*>(src + (size_t)q * g.ys);
        const uint2 p8 = *reinterpret_cast<const uint2*>(&S.pred[by + q][bx]);
        int d[8], y[8];
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const uint32_t sw = x < 4 ? s8.x : s8.y, pw = x < 4 ? p8.x : p8.y;
            d[x] = (int)((sw >> (8 * (x & 3))) & 255) - (int)((pw >> (8 * (x & 3))) & 255);
        }
        fdct8_1d(d, y);
#pragma unroll
        for (int x = 0; x < 8; x++) T[8 * q + ((x + q) & 7)] = y[x];
    }
    __syncwarp();
    const int qbits = 16 + qp / 6, f = (1 << qbits) / 6, rem = qp % 6, sh = qp / 6;
    int nz = 0, cnt4 = 0;   // cnt4: four 8-bit counters, one per interleaved 4x4 block (CAVLC)
    int16_t* lvp = S.lv8;
    {
        int in[8], w[8];
#pragma unroll
        for (int r = 0; r < 8; r++) in[r] = T[8 * r + ((q + r) & 7)];
        fdct8_1d(in, w);
        // quantise; then the decimation rule (vcp_algo.h: vcp_decimate_score over the block's levels in scan order, the
        // macroblock total over the four blocks) decides whether the block is coded at all
        int lq[8];
        unsigned long long m = 0ull;
        bool bigl = false;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int i = 8 * r + q;
            const int l = vcp_quant1(w[r], TT.mf[rem][TT.cls[i]], f, qbits);
            lq[r] = l;
            if (l) m |= 1ull << TT.izz[i];
            bigl |= l > 1 || l < -1;
        }
        m |= __shfl_xor_sync(0xffffffffu, m, 1); m |= __shfl_xor_sync(0xffffffffu, m, 2); m |= __shfl_xor_sync(0xffffffffu, m, 4);
        const int big = (__ballot_sync(0xffffffffu, bigl) >> (8 * k)) & 0xff;
        const int sc = vcp_decimate_score(m, big, 1);
        int tot = sc + __shfl_xor_sync(0xffffffffu, sc, 8);
        tot += __shfl_xor_sync(0xffffffffu, tot, 16);
        const bool drop = g.effort > 0 && vcp_decimate_zero(sc, tot) != 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int i = 8 * r + q;
            const int cls = TT.cls[i], zz = TT.izz[i];
            const int l = drop ? 0 : lq[r];
            nz += l != 0;
            cnt4 += (l != 0) << (8 * (zz & 3));
            lvp[g.cabac ? k * 64 + zz : (k * 4 + (zz & 3)) * 16 + (zz >> 2)] = (int16_t)l;
            const int ls = 16 * TT.v[rem][cls];
            T[8 * r + ((q + r) & 7)] = qp >= 36 ? (l * ls) << (sh - 6) : (l * ls + (1 << (5 - sh))) >> (6 - sh);
        }
    }
    // totals of the 8x8 block over its eight lanes
    nz += __shfl_xor_sync(0xffffffffu, nz, 1); nz += __shfl_xor_sync(0xffffffffu, nz, 2); nz += __shfl_xor_sync(0xffffffffu, nz, 4);
    cnt4 += __shfl_xor_sync(0xffffffffu, cnt4, 1); cnt4 += __shfl_xor_sync(0xffffffffu, cnt4, 2); cnt4 += __shfl_xor_sync(0xffffffffu, cnt4, 4);
    if (q < 4) {
        // nnz of luma4x4BlkIdx 4k+q: CABAC = coefficients of the whole 8x8; CAVLC = of its interleaved
        // 4x4, with bit 7 flagging "the 8x8 holds coefficients" for the deblocking strength
        const int blk = 4 * k + q;
        const int v = g.cabac ? (nz > 255 ? 255 : nz) : (((cnt4 >> (8 * q)) & 255) | (nz ? 0x80 : 0));
        b.nnz[((size_t)gi * g.nmb + mbi) * 24 + blk_y4(blk) * 4 + blk_x4(blk)] = (uint8_t)v;
    }
    __syncwarp();
    if (nz) {   // inverse: row q (uniform over the eight lanes of the block)
        int in[8], o[8];
#pragma unroll
        for (int x = 0; x < 8; x++) in[x] = T[8 * q + ((x + q) & 7)];
        idct8_1d(in, o);
#pragma unroll
        for (int x = 0; x < 8; x++) T[8 * q + ((x + q) & 7)] = o[x];
    }
    __syncwarp();
    if (nz) {   // inverse: column q, then reconstruct in place of the prediction
        int in[8], o[8];
#pragma unroll
        for (int r = 0; r < 8; r++) in[r] = T[8 * r + ((q + r) & 7)];
        idct8_1d(in, o);
#pragma unroll
        for (int r = 0; r < 8; r++)
            S.pred[by + r][bx + q] = (uint8_t)vcp_clip255((int)S.pred[by + r][bx + q] + ((o[r] + 32) >> 6));
    }
    __syncwarp();
    // levels: 512 bytes per macroblock, one 16-byte vector per lane
    reinterpret_cast<uint4*>(b.levels + ((size_t)gi * g.nmb + mbi) * VCP_LV_STRIDE + VCP_LV_LUMA)[lane] = reinterpret_cast<const uint4*>(S.lv8)[lane];
    const uint32_t coded = __ballot_sync(0xffffffffu, nz > 0);
    return ((coded & 0x000000ffu) ? 1u : 0u) | ((coded & 0x0000ff00u) ? 2u : 0u) | ((coded & 0x00ff0000u) ? 4u : 0u) | ((coded & 0xff000000u) ? 8u : 0u);
}

__device__ __forceinline__ void store_block_recon(uint8_t* dst, int stride, const uint32_t pred[4], const int r[16]) {
#pragma unroll
    for (int y = 0; y < 4; y++)
        *reinterpret_cast<uint32_t*>(dst + (size_t)y * stride) =
            vcp_recon4(pred[y], r[4 * y], r[4 * y + 1], r[4 * y + 2], r[4 * y + 3]);
}

// Transform/quantise/reconstruct one macroblock given per-lane predictions.
//  lanes 0..15 : luma block b=lane (pred rows in predw)
//  lanes 16..23: chroma block (plane=(lane-16)>>2, blk=lane&3)
// intra16: luma DC goes through the 4x4 Hadamard (dcbuf = 16 ints of shared scratch).
__device__ __forceinline__ void mb_transform(const VcpGeom& g, const VcpBufs& b, int n, int slot, int gi, int mbi,
                                             int mx, int my, int qp, bool intra16, const uint32_t predw[4],
                                             int* dcbuf, int lane, uint32_t& cbp_out, bool luma_off = false, bool decimate = false) {
    const int qpc = vcp_chroma_qp[vcp_clip3(0, 51, qp)];
    const bool is_luma = lane < 16 && !luma_off, is_chroma = lane >= 16 && lane < 24;
    const int pl = (lane - 16) >> 2, cb = lane & 3;
    int bx = 0, by = 0;
    const uint8_t* src = nullptr;
    uint8_t* dst = nullptr;
    int sstride = 0;
    if (is_luma) {
        bx = blk_x4(lane) * 4; by = blk_y4(lane) * 4;
        src = b.src_y + (size_t)n * g.ysize + g.yoff + (size_t)(16 * my + by) * g.ys + 16 * mx + bx;
        dst = vcp_rec_luma(b, g, slot) + g.yoff + (size_t)(16 * my + by) * g.ys + 16 * mx + bx;
        sstride = g.ys;
    } else if (is_chroma) {
        bx = (cb & 1) * 4; by = (cb >> 1) * 4;
        const size_t o = g.coff + (size_t)(8 * my + by) * g.cs + 8 * mx + bx;
        src = (pl ? b.src_v : b.src_u) + (size_t)n * g.csize + o;
        dst = (pl ? b.rec_v : b.rec_u) + (size_t)slot * g.csize + o;
        sstride = g.cs;
    }
    int w[16], c[16], lv[16], r[16];
    int nz = 0;
    if (is_luma || is_chroma) {
        int d[16];
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const uint32_t s4 = ld_u32(src + (size_t)y * sstride), p4 = predw[y];
#pragma unroll
            for (int x = 0; x < 4; x++) d[4 * y + x] = (int)((s4 >> (8 * x)) & 255) - (int)((p4 >> (8 * x)) & 255);
        }
        vcp_fdct4(d, w);
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = 0;
    }
    const bool ac_only = is_chroma || (is_luma && intra16);
    if (is_luma || is_chroma) nz = vcp_quant_dequant4x4(w, is_luma ? qp : qpc, intra16, ac_only ? 1 : 0, lv, c);
    if (decimate) {   // inter macroblocks (warp-uniform): vcp_algo.h, vcp_decimate_score
        int sc = 0;
        if (is_luma) {
            unsigned long long m = 0ull;
            bool big = false;
#pragma unroll
            for (int i = 0; i < 16; i++) { if (lv[i]) m |= 1ull << i; big |= lv[i] > 1 || lv[i] < -1; }
            sc = vcp_decimate_score(m, big, 0);
        }
        int g8 = sc + __shfl_xor_sync(0xffffffffu, sc, 1);
        g8 += __shfl_xor_sync(0xffffffffu, g8, 2);                 // the four blocks of this lane's 8x8 group
        int tot = g8 + __shfl_xor_sync(0xffffffffu, g8, 4);
        tot += __shfl_xor_sync(0xffffffffu, tot, 8);               // lanes 0..15: the macroblock
        if (is_luma && vcp_decimate_zero(g8, tot)) {
#pragma unroll
            for (int i = 0; i < 16; i++) { lv[i] = 0; c[i] = 0; }
            nz = 0;
        }
    }

    // chroma DC (lanes 16..23; all lanes execute the shuffles)
    {
        int deq = 0;
        const int l = vcp_chroma_dc(is_chroma ? w[0] : 0, qpc, intra16, lane, deq);
        if (is_chroma) {
            c[0] = deq;
            b.levels[((size_t)gi * g.nmb + mbi) * VCP_LV_STRIDE + VCP_LV_CHROMA_DC + pl * 4 + cb] = (int16_t)l;
        }
        const uint32_t dcnz = __ballot_sync(0xffffffffu, is_chroma && l != 0);
        const uint32_t acnz = __ballot_sync(0xffffffffu, is_chroma && nz != 0);
        const uint32_t cbpc = acnz ? 2u : (dcnz ? 1u : 0u);
        cbp_out = cbpc << 4;
        // coded_block_flag of the two chroma DC blocks (CABAC context selection of the neighbours)
        cbp_out |= ((dcnz & 0x000f0000u) ? 0x200u : 0u) | ((dcnz & 0x00f00000u) ? 0x400u : 0u);
        if (is_chroma && !acnz) nz = 0;
    }
    // luma DC for Intra16x16: 4x4 Hadamard over the 16 block DCs (8.5.10 restated forward)
    if (intra16) {
        const int ras = blk_ras(lane & 15);
        // Hadamard sign H[r][c] (rows ++++, ++--, +--+, +-+-): bit r*4+c set => -1
        constexpr uint32_t HS = 0xA6C0u;
        const int hi = (lane >> 2) & 3, hj = lane & 3;
        if (is_luma) dcbuf[ras] = w[0];
        __syncwarp();
        int lvl = 0;
        if (is_luma) {
            int sum = 0;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int sg = (((HS >> (hi * 4 + a)) ^ (HS >> (hj * 4 + q))) & 1) ? -1 : 1;
                    sum += sg * dcbuf[a * 4 + q];
                }
            const int qbits = 15 + qp / 6, f = (1 << qbits) / 3;
            lvl = vcp_quant1(sum >> 1, vcp_quant_mf[qp % 6][0], 2 * f, qbits + 1);
        }
        __syncwarp();
        if (__ballot_sync(0xffffffffu, is_luma && lvl != 0)) cbp_out |= 0x100u;   // coded_block_flag of the luma DC block
        if (is_luma) dcbuf[lane] = lvl;  // raster
        __syncwarp();
        int dq = 0;
        if (is_luma) {
            int sum = 0;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int sg = (((HS >> (hi * 4 + a)) ^ (HS >> (hj * 4 + q))) & 1) ? -1 : 1;
                    sum += sg * dcbuf[a * 4 + q];
                }
            const int ls = 16 * vcp_dequant_v[qp % 6][0];
            dq = qp >= 36 ? (sum * ls) << (qp / 6 - 6) : (sum * ls + (1 << (5 - qp / 6))) >> (6 - qp / 6);
            // zig-zag position of raster index `lane`
            const int zpos = (int)((0xFEA9DB83C7426510ull >> (4 * lane)) & 15);
            b.levels[((size_t)gi * g.nmb + mbi) * VCP_LV_STRIDE + VCP_LV_LUMA_DC + zpos] = (int16_t)lvl;
        }
        __syncwarp();
        if (is_luma) dcbuf[16 + lane] = dq;  // raster dequantised DCs
        __syncwarp();
        if (is_luma) c[0] = dcbuf[16 + ras];
    }
    // coded block pattern (luma)
    {
        const uint32_t nzm = __ballot_sync(0xffffffffu, is_luma && nz != 0);
        uint32_t cbpl;
        if (intra16) { cbpl = nzm ? 15u : 0u; }
        else cbpl = ((nzm & 0x000fu) ? 1u : 0u) | ((nzm & 0x00f0u) ? 2u : 0u) | ((nzm & 0x0f00u) ? 4u : 0u) | ((nzm & 0xf000u) ? 8u : 0u);
        cbp_out |= cbpl;
    }
    // write levels, nnz, recon
    if (is_luma || is_chroma) {
        int16_t* lvp = b.levels + ((size_t)gi * g.nmb + mbi) * VCP_LV_STRIDE +
                       (is_luma ? VCP_LV_LUMA + lane * 16 : VCP_LV_CHROMA_AC + (pl * 4 + cb) * 16);
        vcp_store_levels16(lvp, lv);
        const int ni = is_luma ? blk_y4(lane) * 4 + blk_x4(lane) : 16 + pl * 4 + cb;
        b.nnz[((size_t)gi * g.nmb + mbi) * 24 + ni] = (uint8_t)nz;
        vcp_idct4(c, r);
        store_block_recon(dst, sstride, predw, r);
    }
}

// ---- P macroblocks ---------------------------------------------------------------------------
constexpr int PR_WARPS = 4;

__global__ void __launch_bounds__(PR_WARPS * 32, 8) p_recon_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ McScratch scr[PR_WARPS];
    __shared__ uint8_t cpred[PR_WARPS][2][8][8];
    __shared__ T8Tables t8t;
    if (g.t8x8) {
        if (threadIdx.x < 64) { t8t.cls[threadIdx.x] = vcp_coef8_class[threadIdx.x]; t8t.izz[threadIdx.x] = vcp_izigzag8x8[threadIdx.x]; }
        if (threadIdx.x < 36) { t8t.mf[threadIdx.x / 6][threadIdx.x % 6] = vcp_quant8_mf[threadIdx.x / 6][threadIdx.x % 6]; t8t.v[threadIdx.x / 6][threadIdx.x % 6] = vcp_dequant8_v[threadIdx.x / 6][threadIdx.x % 6]; }
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * PR_WARPS + warp;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    const int n = vcp_frame_of(s, gi);
    const int slot = vcp_rec_slot(s, gi, s.t), rslot = vcp_rec_slot(s, gi, s.t - 1);
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    const int qp = b.qp[n];
    const short2 mv = b.mv[(size_t)gi * g.nmb + mbi];
    if (b.mbtype[(size_t)gi * g.nmb + mbi] == VCP_MB_I16) return;   // decided intra by the refine: i_fix_kernel codes it

    // luma prediction into shared memory
    McScratch& S = scr[warp];
    {
        const int row = lane >> 1, hx = (lane & 1) * 8;
        const uint8_t* blk = vcp_rec_luma(b, g, rslot) + g.yoff + (ptrdiff_t)(16 * my + row) * g.ys + 16 * mx + hx;
        const uint2 p8 = hpel_fetch8(blk, mv.x, mv.y, g.ys, g.ysize);
        *reinterpret_cast<uint2*>(&S.pred[row][hx]) = p8;
    }
    // chroma prediction: lane -> plane, row, 4 px
    {
        const int pl = lane >> 4, row = (lane >> 1) & 7, hx = (lane & 1) * 4;
        const int ix = mv.x >> 3, iy = mv.y >> 3, dx = mv.x & 7, dy = mv.y & 7;
        const uint8_t* cr = (pl ? b.rec_v : b.rec_u) + (size_t)rslot * g.csize + g.coff +
                            (ptrdiff_t)(8 * my + iy + row) * g.cs + 8 * mx + ix + hx;
        const uint2 r0 = ld8_unaligned(cr), r1 = ld8_unaligned(cr + g.cs);
        uint32_t outw = 0;
#pragma unroll
        for (int x = 0; x < 4; x++) {
            const int A = (x < 4 ? (r0.x >> (8 * x)) : 0) & 255;
            const int Bv = x < 3 ? (r0.x >> (8 * (x + 1))) & 255 : r0.y & 255;
            const int Cc = (r1.x >> (8 * x)) & 255;
            const int D = x < 3 ? (r1.x >> (8 * (x + 1))) & 255 : r1.y & 255;
            const int v = ((8 - dx) * (8 - dy) * A + dx * (8 - dy) * Bv + (8 - dx) * dy * Cc + dx * dy * D + 32) >> 6;
            outw |= (uint32_t)v << (8 * x);
        }
        *reinterpret_cast<uint32_t*>(&cpred[warp][pl][row][hx]) = outw;
    }
    __syncwarp();
    uint32_t predw[4] = {0, 0, 0, 0};
    if (lane < 16) {
        const int bx = blk_x4(lane) * 4, by = blk_y4(lane) * 4;
#pragma unroll
        for (int y = 0; y < 4; y++) predw[y] = *reinterpret_cast<const uint32_t*>(&S.pred[by + y][bx]);
    } else if (lane < 24) {
        const int pl = (lane - 16) >> 2, cb = lane & 3, bx = (cb & 1) * 4, by = (cb >> 1) * 4;
#pragma unroll
        for (int y = 0; y < 4; y++) predw[y] = *reinterpret_cast<const uint32_t*>(&cpred[warp][pl][by + y][bx]);
    }
    uint32_t cbp;
    bool use8 = false;
    if (g.t8x8) {
        const uint8_t* sb = b.src_y + (size_t)n * g.ysize + g.yoff + (size_t)(16 * my + blk_y4(lane & 15) * 4) * g.ys + 16 * mx + blk_x4(lane & 15) * 4;
        use8 = prefer_8x8(sb, g.ys, predw, lane);
    }
    if (use8) {
        const uint32_t cbpl = luma8x8_transform(g, b, S, t8t, n, gi, mbi, mx, my, qp, lane);
        if (lane < 16) {   // the reconstruction sits where the prediction was: write this lane's 4x4 block
            const int bx = blk_x4(lane) * 4, by = blk_y4(lane) * 4;
            uint8_t* dst = vcp_rec_luma(b, g, slot) + g.yoff + (size_t)(16 * my + by) * g.ys + 16 * mx + bx;
#pragma unroll
            for (int y = 0; y < 4; y++) *reinterpret_cast<uint32_t*>(dst + (size_t)y * g.ys) = *reinterpret_cast<const uint32_t*>(&S.pred[by + y][bx]);
        }
        mb_transform(g, b, n, slot, gi, mbi, mx, my, qp, false, predw, nullptr, lane, cbp, true);   // chroma only
        cbp = (cbp & ~15u) | cbpl;
        use8 = cbpl != 0;   // transform_size_8x8_flag is only transmitted (else inferred 0) with coded luma
    } else {
        mb_transform(g, b, n, slot, gi, mbi, mx, my, qp, false, predw, nullptr, lane, cbp, false, g.effort > 0);
    }
    if (lane == 0) {
        b.cbp[(size_t)gi * g.nmb + mbi] = (uint8_t)cbp;
        b.mbtype[(size_t)gi * g.nmb + mbi] = VCP_MB_P16;
        // DC coded_block_flags in bits 4..6, transform_size_8x8_flag in bit 7
        b.modes[(size_t)gi * g.nmb + mbi] = (uint8_t)(((cbp >> 8) << 4) | (use8 ? 0x80 : 0));
    }
}

// ---- Intra16x16 macroblocks (wavefront) ----------------------------------------------------------
constexpr int IR_WARPS = 16;

struct __align__(16) IScratch {
    uint8_t top[24];   // [0..3] unused pad, top[4+x] x=-1..16 -> index x+4 (x=-1 at 3)
    uint8_t left[16];
    uint8_t ctop[2][12];  // index x+4, x=-1..7
    uint8_t cleft[2][8];
    uint8_t pred[16][16];
    uint8_t cpred[2][8][8];
    int dcbuf[32];
};

__device__ __forceinline__ uint32_t splat4(int v) { return (uint32_t)v * 0x01010101u; }

__device__ void i16_encode_mb(const VcpGeom& g, const VcpBufs& b, IScratch& S, int n, int slot, int gi, int mx, int my,
                              int row0, int qp, int lane) {
    const int mbi = my * g.mbw + mx;
    const bool aL = mx > 0, aT = my > row0;
    uint8_t* ry = vcp_rec_luma(b, g, slot) + g.yoff + (size_t)(16 * my) * g.ys + 16 * mx;
    uint8_t* ru = b.rec_u + (size_t)slot * g.csize + g.coff + (size_t)(8 * my) * g.cs + 8 * mx;
    uint8_t* rv = b.rec_v + (size_t)slot * g.csize + g.coff + (size_t)(8 * my) * g.cs + 8 * mx;
    // neighbours (unfiltered reconstruction of this picture)
    if (lane < 17) S.top[3 + lane] = (aT && (lane > 0 || aL)) ? ry[-(ptrdiff_t)g.ys + lane - 1] : 128;
    if (lane < 16) S.left[lane] = aL ? ry[(ptrdiff_t)lane * g.ys - 1] : 128;
    if (lane < 18) {
        const int pl = lane / 9, x = lane % 9 - 1;
        const uint8_t* rc = pl ? rv : ru;
        S.ctop[pl][4 + x] = (aT && (x >= 0 || aL)) ? rc[-(ptrdiff_t)g.cs + x] : 128;
    }
    if (lane >= 16) {
        const int pl = (lane - 16) >> 3, y = lane & 7;
        S.cleft[pl][y] = aL ? (pl ? rv : ru)[(ptrdiff_t)y * g.cs - 1] : 128;
    }
    __syncwarp();
    // ---- luma mode decision: lane owns row = lane>>1, 8 px at hx
    const int row = lane >> 1, hx = (lane & 1) * 8;
    const uint8_t* sy = b.src_y + (size_t)n * g.ysize + g.yoff + (size_t)(16 * my + row) * g.ys + 16 * mx + hx;
    const uint2 c8 = *reinterpret_cast<const uint2*>(sy);
    uint32_t best = 0xffffffffu;
    uint2 bestp = make_uint2(0, 0);
    // DC value
    int dcv;
    {
        int st = 0, sl = 0;
        for (int i = 0; i < 16; i++) { st += S.top[4 + i]; sl += S.left[i]; }
        dcv = (aT && aL) ? (st + sl + 16) >> 5 : aT ? (st + 8) >> 4 : aL ? (sl + 8) >> 4 : 128;
    }
    // plane parameters
    int pa = 0, pb = 0, pc = 0;
    if (aT && aL) {
        int H = 0, V = 0;
        for (int i = 0; i < 8; i++) {
            H += (i + 1) * (S.top[4 + 8 + i] - S.top[4 + 6 - i]);
            const int lo = 6 - i;
            V += (i + 1) * (S.left[8 + i] - (lo >= 0 ? S.left[lo] : S.top[3]));
        }
        pa = 16 * (S.left[15] + S.top[4 + 15]); pb = (5 * H + 32) >> 6; pc = (5 * V + 32) >> 6;
    }
#pragma unroll 1
    for (int mode = 0; mode < 4; mode++) {
        if ((mode == 0 && !aT) || (mode == 1 && !aL) || (mode == 3 && !(aT && aL))) continue;
        uint2 p;
        if (mode == 0) p = make_uint2(*reinterpret_cast<const uint32_t*>(&S.top[4 + hx]), *reinterpret_cast<const uint32_t*>(&S.top[8 + hx]));
        else if (mode == 1) p = make_uint2(splat4(S.left[row]), splat4(S.left[row]));
        else if (mode == 2) p = make_uint2(splat4(dcv), splat4(dcv));
        else {
            uint32_t w0 = 0, w1 = 0;
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const uint32_t v = (uint32_t)vcp_clip255((pa + pb * (hx + x - 7) + pc * (row - 7) + 16) >> 5);
                if (x < 4) w0 |= v << (8 * x); else w1 |= v << (8 * (x - 4));
            }
            p = make_uint2(w0, w1);
        }
        const int sad = warp_sum((int)sad4(p.y, c8.y, sad4(p.x, c8.x, 0)));
        const uint32_t key = ((uint32_t)sad << 2) | (uint32_t)mode;
        if (key < best) { best = key; bestp = p; }
    }
    const int i16mode = (int)(best & 3);
    *reinterpret_cast<uint2*>(&S.pred[row][hx]) = bestp;

    // ---- chroma mode decision: lane -> plane = lane>>4, row = (lane>>1)&7, 4 px at cx
    const int cpl = lane >> 4, crow = (lane >> 1) & 7, cx = (lane & 1) * 4;
    const uint8_t* sc = (cpl ? b.src_v : b.src_u) + (size_t)n * g.csize + g.coff + (size_t)(8 * my + crow) * g.cs + 8 * mx + cx;
    const uint32_t cc4 = ld_u32(sc);
    best = 0xffffffffu;
    uint32_t bestc = 0;
    // DC of this lane's 4x4 block
    int cdc;
    {
        const int bxx = cx, byy = crow & 4;
        int st = 0, sl = 0;
        for (int i = 0; i < 4; i++) { st += S.ctop[cpl][4 + bxx + i]; sl += S.cleft[cpl][byy + i]; }
        const int blk = (byy >> 2) * 2 + (bxx >> 2);
        if (blk == 0 || blk == 3) cdc = (aT && aL) ? (st + sl + 4) >> 3 : aT ? (st + 2) >> 2 : aL ? (sl + 2) >> 2 : 128;
        else if (blk == 1) cdc = aT ? (st + 2) >> 2 : aL ? (sl + 2) >> 2 : 128;
        else cdc = aL ? (sl + 2) >> 2 : aT ? (st + 2) >> 2 : 128;
    }
    int ca = 0, cbb = 0, ccc = 0;
    if (aT && aL) {
        int H = 0, V = 0;
        for (int i = 0; i < 4; i++) {
            H += (i + 1) * (S.ctop[cpl][4 + 4 + i] - S.ctop[cpl][4 + 2 - i]);
            const int lo = 2 - i;
            V += (i + 1) * (S.cleft[cpl][4 + i] - (lo >= 0 ? S.cleft[cpl][lo] : S.ctop[cpl][3]));
        }
        ca = 16 * (S.cleft[cpl][7] + S.ctop[cpl][4 + 7]); cbb = (34 * H + 32) >> 6; ccc = (34 * V + 32) >> 6;
    }
#pragma unroll 1
    for (int mode = 0; mode < 4; mode++) {
        if ((mode == 1 && !aL) || (mode == 2 && !aT) || (mode == 3 && !(aT && aL))) continue;
        uint32_t p;
        if (mode == 0) p = splat4(cdc);
        else if (mode == 1) p = splat4(S.cleft[cpl][crow]);
        else if (mode == 2) p = *reinterpret_cast<const uint32_t*>(&S.ctop[cpl][4 + cx]);
        else {
            p = 0;
#pragma unroll
            for (int x = 0; x < 4; x++)
                p |= (uint32_t)vcp_clip255((ca + cbb * (cx + x - 3) + ccc * (crow - 3) + 16) >> 5) << (8 * x);
        }
        const int sad = warp_sum((int)sad4(p, cc4, 0));
        const uint32_t key = ((uint32_t)sad << 2) | (uint32_t)mode;
        if (key < best) { best = key; bestc = p; }
    }
    const int cmode = (int)(best & 3);
    *reinterpret_cast<uint32_t*>(&S.cpred[cpl][crow][cx]) = bestc;
    __syncwarp();

    uint32_t predw[4] = {0, 0, 0, 0};
    if (lane < 16) {
        const int bx = blk_x4(lane) * 4, by = blk_y4(lane) * 4;
#pragma unroll
        for (int y = 0; y < 4; y++) predw[y] = *reinterpret_cast<const uint32_t*>(&S.pred[by + y][bx]);
    } else if (lane < 24) {
        const int pl = (lane - 16) >> 2, cb = lane & 3, bx = (cb & 1) * 4, by = (cb >> 1) * 4;
#pragma unroll
        for (int y = 0; y < 4; y++) predw[y] = *reinterpret_cast<const uint32_t*>(&S.cpred[pl][by + y][bx]);
    }
    uint32_t cbp;
    mb_transform(g, b, n, slot, gi, mbi, mx, my, qp, true, predw, S.dcbuf, lane, cbp);
    if (lane == 0) {
        const size_t o = (size_t)gi * g.nmb + mbi;
        b.cbp[o] = (uint8_t)cbp;
        b.mbtype[o] = VCP_MB_I16;
        b.modes[o] = (uint8_t)(i16mode | (cmode << 2) | ((cbp >> 8) << 4));
        b.mv[o] = make_short2(0, 0);
        b.mvd[o] = make_short2(0, 0);
    }
}

// grid: x = slice, y = GOP
__global__ void __launch_bounds__(IR_WARPS * 32) i_recon_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ IScratch scr[IR_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = blockIdx.x, gi = blockIdx.y + s.g0;
    const int n = vcp_frame_of(s, gi);
    const int slot = vcp_rec_slot(s, gi, s.t);
    const int qp = b.qp[n];
    const int r0 = vcp_slice_first_row(sl, g.slices, g.mbh);
    const int r1 = sl + 1 < g.slices ? vcp_slice_first_row(sl + 1, g.slices, g.mbh) : g.mbh;
    const int rows = r1 - r0;
    const int ndiag = g.mbw + rows - 1;
    for (int d = 0; d < ndiag; d++) {
        // macroblocks on this anti-diagonal: (mx, r0 + k) with mx = d - k
        const int k0 = d - (g.mbw - 1) > 0 ? d - (g.mbw - 1) : 0;
        const int k1 = d < rows - 1 ? d : rows - 1;
        for (int k = k0 + warp; k <= k1; k += IR_WARPS)
            i16_encode_mb(g, b, scr[warp], n, slot, gi, d - k, r0 + k, r0, qp, lane);
        __syncthreads();
    }
}

// Intra16x16 macroblocks INSIDE P pictures (flagged by the refine).  They predict from the
// reconstruction of their left/top neighbours in the same picture: inter neighbours are complete
// once p_recon_kernel has run, intra neighbours come earlier on the anti-diagonal wavefront.
// grid: x = slice, y = GOP; a slice without flagged macroblocks leaves at once.
__global__ void __launch_bounds__(IR_WARPS * 32) i_fix_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ IScratch scr[IR_WARPS];
    __shared__ int todo;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = blockIdx.x, gi = blockIdx.y + s.g0;
    if (threadIdx.x == 0) { todo = b.icount[(size_t)gi * g.slices + sl]; b.icount[(size_t)gi * g.slices + sl] = 0; }
    __syncthreads();
    if (!todo) return;
    const int n = vcp_frame_of(s, gi);
    const int slot = vcp_rec_slot(s, gi, s.t);
    const int qp = b.qp[n];
    const int r0 = vcp_slice_first_row(sl, g.slices, g.mbh);
    const int r1 = sl + 1 < g.slices ? vcp_slice_first_row(sl + 1, g.slices, g.mbh) : g.mbh;
    const int rows = r1 - r0;
    const int ndiag = g.mbw + rows - 1;
    // which anti-diagonals hold intra macroblocks at all: only those cost a barrier
    __shared__ uint32_t dmask[32];   // up to 1024 diagonals
    if (threadIdx.x < 32) dmask[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < rows * g.mbw; i += blockDim.x)
        if (b.mbtype[(size_t)gi * g.nmb + r0 * g.mbw + i] == VCP_MB_I16) {
            const int d = i % g.mbw + i / g.mbw;
            atomicOr(&dmask[(d >> 5) & 31], 1u << (d & 31));
        }
    __syncthreads();
    for (int d = 0; d < ndiag; d++) {
        if (!((dmask[(d >> 5) & 31] >> (d & 31)) & 1)) continue;
        const int k0 = d - (g.mbw - 1) > 0 ? d - (g.mbw - 1) : 0;
        const int k1 = d < rows - 1 ? d : rows - 1;
        for (int k = k0 + warp; k <= k1; k += IR_WARPS) {
            const int mx = d - k, my = r0 + k;
            if (b.mbtype[(size_t)gi * g.nmb + my * g.mbw + mx] == VCP_MB_I16)
                i16_encode_mb(g, b, scr[warp], n, slot, gi, mx, my, r0, qp, lane);
        }
        __syncthreads();   // reconstruction of this diagonal is visible to the next
    }
}

}  // namespace

void vcp_launch_i_fix(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid(g.slices, s.ngop);
    i_fix_kernel<<<grid, IR_WARPS * 32, 0, st>>>(g, b, s);
}

void vcp_launch_p_recon(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + PR_WARPS - 1) / PR_WARPS, s.ngop);
    p_recon_kernel<<<grid, PR_WARPS * 32, 0, st>>>(g, b, s);
}

void vcp_launch_i_recon(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid(g.slices, s.ngop);
    i_recon_kernel<<<grid, IR_WARPS * 32, 0, st>>>(g, b, s);
}
