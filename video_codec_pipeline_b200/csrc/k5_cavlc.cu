// K5 — CAVLC entropy coding (H.264 7.3.4-7.3.5, 9.2) and NAL encapsulation (7.4.1, Annex B).
//
// CAVLC has no adaptive state: the bits of a macroblock depend only on its own levels and on
// neighbour *counts* (nC) and vectors that are all known once K3/K4 have run.  So instead of
// one sequential writer per slice, the bitstream is produced macroblock-parallel in passes:
//   cavlc_mb_kernel<false> : one warp per macroblock, one lane per syntax group (header, Intra16x16 DC,
//                            16 luma blocks, 2 chroma DC, 8 chroma AC) -> bits per macroblock.
//   cavlc_scan_kernel      : one CTA per slice: exclusive prefix sum of macroblock bits, slice
//                            header (7.3.3) and trailing skip run / RBSP trailing bits, zeroing of
//                            exactly the bytes the slice will occupy.
//   cavlc_mb_kernel<true>  : same coder again, now OR-ing each lane's code words into the slice
//                            RBSP at its absolute bit offset.
//   nal_pack_kernel        : start code + NAL header + emulation prevention (00 00 0x -> 00 00 03
//                            0x) done as a parallel scan, appended to the output arena.
//
// Replaces x264's cavlc writer + libavformat's Annex-B framing inside the ffmpeg child
// (/root/reference/cmd/consumer.go:376-382).  Output bytes are identical to
// oracle/h264_oracle.c (write_slice_header, write_slice_data, cavlc_block, nal_write).
#include "vcp_dev.cuh"

#define VCP_TAB static __device__ const
#include "h264_tables.h"
#include "vcp_entropy.cuh"

namespace {

constexpr int CV_WARPS = 4;
constexpr int LANE_WORDS = 17;

// ---- per-lane bit sink ---------------------------------------------------------------------------
template <bool WRITE>
struct BitSink {
    uint32_t* words;  // LANE_WORDS words of shared memory (WRITE only)
    uint64_t acc;
    int nacc;   // bits in acc (< 32 after flush)
    int nbits;  // total
    __device__ __forceinline__ void init(uint32_t* w) { words = w; acc = 0; nacc = 0; nbits = 0; }
    __device__ __forceinline__ void put(int n, uint32_t v) {  // n <= 32
        nbits += n;
        if (WRITE) {
            acc = (acc << n) | (uint64_t)v;
            nacc += n;
            if (nacc >= 32) {
                words[(nbits - nacc) >> 5] = (uint32_t)(acc >> (nacc - 32));
                nacc -= 32;
            }
        }
    }
    __device__ __forceinline__ void ue(uint32_t k) {
        const uint32_t x = k + 1;
        const int n = 31 - __clz(x);
        put(n, 0);
        put(n + 1, x);
    }
    __device__ __forceinline__ void se(int v) { ue(v <= 0 ? (uint32_t)(-2 * v) : (uint32_t)(2 * v - 1)); }
    __device__ __forceinline__ void finish() {
        if (WRITE && nacc > 0) words[(nbits - nacc) >> 5] = (uint32_t)(acc << (32 - nacc));
    }
};

// residual_block_cavlc for levels c[0..n-1] (scan order) held in shared memory
template <bool WRITE>
__device__ __forceinline__ void cavlc_block(BitSink<WRITE>& bs, const int16_t* c, int n, int nC) {
    uint32_t mask = 0;
    for (int i = 0; i < n; i++) mask |= (c[i] != 0 ? 1u : 0u) << i;
    const int total = __popc(mask);
    // trailing ones: up to three +-1 from the high-frequency end
    int t1 = 0;
    {
        uint32_t m = mask;
        while (m && t1 < 3) {
            const int i = 31 - __clz(m);
            const int v = c[i];
            if (v != 1 && v != -1) break;
            t1++;
            m &= ~(1u << i);
        }
    }
    if (nC == -1) bs.put(vcp_chroma_dc_coeff_token_len[4 * total + t1], vcp_chroma_dc_coeff_token_bits[4 * total + t1]);
    else {
        const int tab = nC < 2 ? 0 : nC < 4 ? 1 : nC < 8 ? 2 : 3;
        bs.put(vcp_coeff_token_len[tab][4 * total + t1], vcp_coeff_token_bits[tab][4 * total + t1]);
    }
    if (!total) return;
    const int last = 31 - __clz(mask);
    uint32_t m = mask;
    int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
    for (int k = 0; k < total; k++) {
        const int i = 31 - __clz(m);
        m &= ~(1u << i);
        const int l = c[i];
        if (k < t1) { bs.put(1, l < 0); continue; }
        int code = l > 0 ? 2 * l - 2 : -2 * l - 1;
        if (k == t1 && t1 < 3) code -= 2;
        if (suffix_len == 0) {
            if (code < 14) bs.put(code + 1, 1);
            else if (code < 30) { bs.put(15, 1); bs.put(4, code - 14); }
            else { bs.put(16, 1); bs.put(12, code - 30); }
        } else {
            const int pre = code >> suffix_len;
            if (pre < 15) { bs.put(pre + 1, 1); bs.put(suffix_len, code & ((1 << suffix_len) - 1)); }
            else { bs.put(16, 1); bs.put(12, code - (15 << suffix_len)); }
        }
        if (suffix_len == 0) suffix_len = 1;
        if (vcp_iabs(l) > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
    }
    const int total_zeros = last + 1 - total;
    if (total < n) {
        if (n == 4) bs.put(vcp_chroma_dc_total_zeros_len[total - 1][total_zeros], vcp_chroma_dc_total_zeros_bits[total - 1][total_zeros]);
        else bs.put(vcp_total_zeros_len[total - 1][total_zeros], vcp_total_zeros_bits[total - 1][total_zeros]);
    }
    // run_before, from the high-frequency end
    int left = total_zeros;
    m = mask;
    int prev = 31 - __clz(m);
    m &= ~(1u << prev);
    while (m && left > 0) {
        const int i = 31 - __clz(m);
        m &= ~(1u << i);
        const int run = prev - i - 1;
        const int zl = left > 7 ? 7 : left;
        bs.put(vcp_run_len[zl - 1][run], vcp_run_bits[zl - 1][run]);
        left -= run;
        prev = i;
    }
}

struct __align__(16) CvScratch {
    int16_t lv[VCP_LV_STRIDE];
    uint32_t words[32][LANE_WORDS];
    uint8_t nnz[3][24];  // cur, left, top
};

// nC of block (bx,by) in units of blocks; plane 0: luma (4x4 grid), 1/2: chroma (2x2 grid)
__device__ __forceinline__ int nnz_ctx(const CvScratch& S, bool aL, bool aT, int bx, int by, int plane) {
    int nA = -1, nB = -1;
    if (plane == 0) {
        if (bx > 0) nA = S.nnz[0][by * 4 + bx - 1]; else if (aL) nA = S.nnz[1][by * 4 + 3];
        if (by > 0) nB = S.nnz[0][(by - 1) * 4 + bx]; else if (aT) nB = S.nnz[2][12 + bx];
    } else {
        const int o = 16 + (plane - 1) * 4;
        if (bx > 0) nA = S.nnz[0][o + by * 2]; else if (aL) nA = S.nnz[1][o + by * 2 + 1];
        if (by > 0) nB = S.nnz[0][o + bx]; else if (aT) nB = S.nnz[2][o + 2 + bx];
    }
    if (nA >= 0 && nB >= 0) return (nA + nB + 1) >> 1;
    return nA >= 0 ? nA : (nB >= 0 ? nB : 0);
}

template <bool WRITE>
__global__ void __launch_bounds__(CV_WARPS * 32) cavlc_mb_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ CvScratch scr[CV_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * CV_WARPS + warp;
    const int gi = blockIdx.y + s.g0;
    if (mbi >= g.nmb) return;
    if (WRITE) {
        int err = lane == 0 ? *b.error_flag : 0;
        err = __shfl_sync(0xffffffffu, err, 0);
        if (err) return;
    }
    const size_t o = (size_t)gi * g.nmb + mbi;
    const int type = b.mbtype[o];
    if (type == VCP_MB_PSKIP) { if (!WRITE && lane == 0) b.mbbits[o] = 0; return; }
    const bool idr = s.t == 0;
    const int mx = mbi % g.mbw, my = mbi / g.mbw;
    const int sl = vcp_row_slice(b, my);
    const int row0 = vcp_row_first(b, my);
    const bool aL = mx > 0, aT = my > row0;
    CvScratch& S = scr[warp];
    // stage levels (816 B = 51 x 16 B) and the nnz of cur / left / top
    for (int i = lane; i < VCP_LV_STRIDE * 2 / 16; i += 32)
        reinterpret_cast<uint4*>(S.lv)[i] = reinterpret_cast<const uint4*>(b.levels + o * VCP_LV_STRIDE)[i];
    if (lane < 18) {
        const int w = lane / 6, c = lane % 6;
        const size_t src = w == 0 ? o : (w == 1 ? o - 1 : o - g.mbw);
        const bool ok = w == 0 || (w == 1 ? aL : aT);
        // bit 7 of a luma count only tells the deblocking filter that the 8x8 block is coded
        reinterpret_cast<uint32_t*>(S.nnz[w])[c] = ok ? (reinterpret_cast<const uint32_t*>(b.nnz + src * 24)[c] & 0x7f7f7f7fu) : 0u;
    }
    // preceding skip run (P slices): cooperative look-back inside the slice
    int skip_run = 0;
    if (!idr) {
        const int first = row0 * g.mbw;
        int pos = mbi - 1;
        while (pos >= first) {
            const int p = pos - lane;
            const bool sk = p >= first && b.mbtype[(size_t)gi * g.nmb + p] == VCP_MB_PSKIP;
            const uint32_t nm = ~__ballot_sync(0xffffffffu, sk);
            if (nm) { skip_run += __ffs(nm) - 1; break; }
            skip_run += 32; pos -= 32;
        }
    }
    __syncwarp();
    const int cbp = b.cbp[o];
    const int cbpl = cbp & 15, cbpc = cbp >> 4;
    BitSink<WRITE> bs;
    bs.init(S.words[lane]);
    if (lane == 0) {
        if (!idr) bs.ue((uint32_t)skip_run);
        if (type == VCP_MB_I16) {
            const int modes = b.modes[o];
            const int t = 1 + (modes & 3) + 4 * cbpc + (cbpl ? 12 : 0);
            bs.ue((uint32_t)(idr ? t : t + 5));
            bs.ue((uint32_t)((modes >> 2) & 3));
            bs.se(0);
        } else {
            const short2 d = b.mvd[o];
            bs.ue(0);
            bs.se(d.x); bs.se(d.y);
            bs.ue(vcp_cbp_to_golomb_inter[cbp]);
            if (g.t8x8 && cbpl) bs.put(1, (uint32_t)(b.modes[o] >> 7));   // transform_size_8x8_flag
            if (cbp) bs.se(0);
        }
    } else if (lane == 1) {
        if (type == VCP_MB_I16) cavlc_block<WRITE>(bs, S.lv + VCP_LV_LUMA_DC, 16, nnz_ctx(S, aL, aT, 0, 0, 0));
    } else if (lane < 18) {
        const int blk = lane - 2;
        if (cbpl & (1 << (blk >> 2))) {
            const int bx = (blk & 1) | ((blk >> 1) & 2), by = ((blk >> 1) & 1) | ((blk >> 2) & 2);
            const int nC = nnz_ctx(S, aL, aT, bx, by, 0);
            const int16_t* lv = S.lv + VCP_LV_LUMA + blk * 16;
            if (type == VCP_MB_I16) cavlc_block<WRITE>(bs, lv + 1, 15, nC); else cavlc_block<WRITE>(bs, lv, 16, nC);
        }
    } else if (lane < 20) {
        if (cbpc) cavlc_block<WRITE>(bs, S.lv + VCP_LV_CHROMA_DC + (lane - 18) * 4, 4, -1);
    } else if (lane < 28) {
        if (cbpc & 2) {
            const int pl = (lane - 20) >> 2, blk = lane & 3;
            cavlc_block<WRITE>(bs, S.lv + VCP_LV_CHROMA_AC + (pl * 4 + blk) * 16 + 1, 15, nnz_ctx(S, aL, aT, blk & 1, blk >> 1, pl + 1));
        }
    }
    bs.finish();
    // exclusive scan of lane bit counts
    int incl = bs.nbits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (!WRITE) {
        if (lane == 0) b.mbbits[o] = (uint32_t)total;
        return;
    }
    // place this lane's words at its absolute bit offset (big-endian bit order)
    uint32_t* dst = reinterpret_cast<uint32_t*>(b.rbsp + ((size_t)gi * g.slices + sl) * b.rbsp_cap);
    const uint32_t pos0 = b.mbbitoff[o] + (uint32_t)(incl - bs.nbits);
    const int nw = (bs.nbits + 31) >> 5;
    for (int k = 0; k < nw; k++) {
        const uint32_t v = S.words[lane][k];
        const uint32_t p = pos0 + 32u * k;
        const uint32_t wi = p >> 5, sh = p & 31;
        const uint32_t hi = v >> sh;
        if (hi) atomicOr(&dst[wi], __byte_perm(hi, 0, 0x0123));
        if (sh) {
            const uint32_t lo = v << (32 - sh);
            if (lo) atomicOr(&dst[wi + 1], __byte_perm(lo, 0, 0x0123));
        }
    }
}

constexpr int SCAN_THREADS = 512;

// grid: x = slice, y = GOP
__global__ void __launch_bounds__(SCAN_THREADS) cavlc_scan_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    __shared__ uint32_t wsum[33];
    __shared__ int last_nonskip;
    const int sl = blockIdx.x, gi = blockIdx.y + s.g0;
    const int n = vcp_frame_of(s, gi);
    const bool idr = s.t == 0;
    const int qp = b.qp[n];
    const int r0 = vcp_slice_first_row(sl, g.slices, g.mbh);
    const int r1 = sl + 1 < g.slices ? vcp_slice_first_row(sl + 1, g.slices, g.mbh) : g.mbh;
    const int first = r0 * g.mbw, count = (r1 - r0) * g.mbw;
    const size_t base = (size_t)gi * g.nmb + first;
    const int gop_index = s.gop0 + n / s.gop;
    const int hdr = slice_header_bits(g, first, idr, s.t, gop_index & 1, qp, nullptr);
    if (threadIdx.x == 0) last_nonskip = -1;
    __syncthreads();
    uint32_t carry = (uint32_t)hdr;
    int my_last = -1;
    for (int i0 = 0; i0 < count; i0 += SCAN_THREADS) {
        const int i = i0 + threadIdx.x;
        const uint32_t v = i < count ? b.mbbits[base + i] : 0;
        if (i < count && b.mbtype[base + i] != VCP_MB_PSKIP) my_last = i;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(v, wsum, tot);
        if (i < count) b.mbbitoff[base + i] = carry + ex;
        carry += tot;
    }
    if (my_last >= 0) atomicMax(&last_nonskip, my_last);
    __syncthreads();
    const int trail_run = idr ? 0 : count - 1 - last_nonskip;
    uint32_t total = carry;
    if (trail_run > 0) total += (uint32_t)vcp_ue_len((unsigned)trail_run);
    total += 1;                        // rbsp_stop_one_bit
    total = (total + 7u) & ~7u;        // alignment zero bits
    const uint32_t bytes = total >> 3;
    uint8_t* dst = b.rbsp + ((size_t)gi * g.slices + sl) * b.rbsp_cap;
    if ((size_t)bytes + 8 > b.rbsp_cap) {
        if (threadIdx.x == 0) { atomicExch(b.error_flag, 1); b.slice_bits[(size_t)gi * g.slices + sl] = 0; }
        return;
    }
    // zero exactly the words this slice will touch
    const uint32_t nwords = (bytes + 3) / 4 + 1;
    for (uint32_t i = threadIdx.x; i < nwords; i += SCAN_THREADS) reinterpret_cast<uint32_t*>(dst)[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        SeqBits w{dst, 0};
        slice_header_bits(g, first, idr, s.t, gop_index & 1, qp, &w);
        w.pos = carry;
        if (trail_run > 0) w.ue((uint32_t)trail_run);
        w.put(1, 1);
        b.slice_bits[(size_t)gi * g.slices + sl] = total;
        atomicAdd(&b.frame_bits[n], total);
    }
}

// grid: x = slice, y = GOP.  NAL = 00 00 00 01 | header | escaped RBSP
__global__ void __launch_bounds__(PACK_THREADS) nal_pack_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    const int sl = blockIdx.x, gi = blockIdx.y + s.g0;
    const int n = vcp_frame_of(s, gi);
    const uint32_t bytes = b.slice_bits[(size_t)gi * g.slices + sl] >> 3;
    const uint8_t* src = b.rbsp + ((size_t)gi * g.slices + sl) * b.rbsp_cap;
    nal_pack_body(g, b, src, bytes, n, sl, s.t == 0);
}

// ---- rate control feedback (one thread per GOP) -------------------------------------------------
__global__ void rc_update_kernel(VcpGeom g, VcpBufs b, VcpStep s) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= s.ngop) return;
    const int gi = s.g0 + k;
    const int n = vcp_frame_of(s, gi);
    int gop_len = s.gop;
    if (gi * s.gop + gop_len > s.nframes) gop_len = s.nframes - gi * s.gop;
    // the QP of picture t+2 (vcp_algo.h); the state is advanced for every picture
    unsigned long long cum = b.rc_cum[gi];
    long long full = b.rc_full[gi];
    const int q = vcp_rc_picture(g.rc_abr, g.rc_qp0, g.rc_qp_nom,
                                 g.rc_abr ? vcp_rc_gop_budget(g.rc_bitrate, g.rc_maxrate, g.fps_num, g.fps_den, gop_len) : 0ull,
                                 g.rc_vbv_rate, g.rc_vbv_buf, b.qp[n], s.t + 1 < gop_len ? b.qp[n + 1] : b.qp[n], s.t == 0,
                                 b.frame_bits[n], s.t, gop_len, &cum, &full);
    b.rc_cum[gi] = cum;
    b.rc_full[gi] = full;
    if (s.t + 2 < gop_len) b.qp[n + 2] = (uint8_t)q;
}

}  // namespace

void vcp_launch_rc_update(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    if (!g.rc_fb) return;
    rc_update_kernel<<<(s.ngop + 63) / 64, 64, 0, st>>>(g, b, s);
}

void vcp_launch_cavlc_count(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + CV_WARPS - 1) / CV_WARPS, s.ngop);
    cavlc_mb_kernel<false><<<grid, CV_WARPS * 32, 0, st>>>(g, b, s);
}
void vcp_launch_cavlc_scan(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid(g.slices, s.ngop);
    cavlc_scan_kernel<<<grid, SCAN_THREADS, 0, st>>>(g, b, s);
}
void vcp_launch_cavlc_write(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid((g.nmb + CV_WARPS - 1) / CV_WARPS, s.ngop);
    cavlc_mb_kernel<true><<<grid, CV_WARPS * 32, 0, st>>>(g, b, s);
}
void vcp_launch_nal_pack(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st) {
    dim3 grid(g.slices, s.ngop);
    nal_pack_kernel<<<grid, PACK_THREADS, 0, st>>>(g, b, s);
}
