// Encoder-side algorithm constants (non-normative choices) shared by the CUDA kernels and
// by the CPU oracle, so that both make the same decisions and emit the same bitstream.
// Everything here is a *choice of this encoder*; the normative tables are in h264_tables.h.
#ifndef VCP_ALGO_H
#define VCP_ALGO_H

#include <stdint.h>

#if defined(__CUDACC__)
#define VCP_HD __host__ __device__ __forceinline__
#else
#define VCP_HD static inline
#endif

// Frame planes carry a replicated border so motion vectors may point outside the picture
// (H.264 unrestricted MVs) without clamping in the inner loops.
#define VCP_PAD 32    // luma border, pixels
#define VCP_PADC 16   // chroma border
#define VCP_PAD1 16   // border of the half-resolution luma plane used by the ME pre-pass

// Motion search geometry (see DESIGN.md "K2"):
//  pre-pass on ORIGINAL frames (all frames in parallel, off the recon dependency chain):
//    L1: half-res 8x8 block per MB, full search +-VCP_ME_R1 (=> +-2*R1 full-pel)
//    L0: full-res 16x16, +-2 around 2*mvL1
//  refine on the RECONSTRUCTED reference (inside the per-frame chain):
//    full-pel +-1 (+ zero and predictor candidates), half-pel 8 pts, quarter-pel 8 pts
#define VCP_ME_R1 12
#define VCP_ME_L1_PEN 2   // cost += PEN * (|dx|+|dy|) at L1 (half-res pixels)
#define VCP_ME_L0_PEN 2   // cost += PEN * (|mvx|+|mvy|) at L0 (full-res pixels)
#define VCP_MV_FP_MAX 27  // |full-pel mv| bound after refine; +-111 in quarter units
// a macroblock whose best full-pel cost is already below this keeps the full-pel vector
// (static / perfectly tracked content: sub-pel refinement cannot pay for itself)
#define VCP_SUBPEL_SKIP_COST 256

// Intra16x16 inside P pictures: chosen when the intra estimate (SAD of the best of V/H/DC
// prediction from ORIGINAL neighbours) beats the inter cost by a margin; intra macroblocks cost
// more side information and cannot be skipped, hence the 25 % + 16 lambda handicap.
VCP_HD int vcp_intra_wins(int intra_sad, int inter_cost, int lam) {
    return intra_sad + (intra_sad >> 2) + 16 * lam < inter_cost;
}

// ---- HEVC sample adaptive offset (luma edge offsets), one decision per 16x16 coding tree block -------------
// sum / cnt: (source - deblocked) accumulated over the samples of one edge category.  Categories 1, 2 (local
// minimum, concave corner) take offsets 0..7, categories 3, 4 (convex corner, local maximum) -7..0 (7.4.9.3.2).
VCP_HD int vcp_sao_offset(int sum, int cnt, int positive) {
    if (cnt == 0) return 0;
    int o = sum >= 0 ? (sum + (cnt >> 1)) / cnt : -((-sum + (cnt >> 1)) / cnt);
    if (positive) return o < 0 ? 0 : o > 7 ? 7 : o;
    return o > 0 ? 0 : o < -7 ? -7 : o;
}
// Cost of coding one class with its four offsets against "off" (cost lam2 for its single bin): distortion change
// sum(cnt o^2 - 2 o sum) plus lam2 per bin (type 2, offsets |o| + 1 each up to 7, class 2).  Returns the class
// cost; off[] receives the offsets.
VCP_HD long long vcp_sao_class_cost(const int sum[4], const int cnt[4], int lam2, int off[4]) {
    long long d = 0;
    int bins = 4;
    for (int k = 0; k < 4; k++) {
        const int o = vcp_sao_offset(sum[k], cnt[k], k < 2);
        off[k] = o;
        d += (long long)cnt[k] * o * o - 2LL * o * sum[k];
        const int a = o < 0 ? -o : o;
        bins += a + (a < 7);
    }
    return d + (long long)lam2 * bins;
}

// Slices per picture when the caller leaves the choice to the encoder (slices == 0).  CAVLC is
// macroblock-parallel, one slice is best.  CABAC is one sequential chain per slice, so pictures are cut
// into slices of about 17 macroblock rows (1080p: 4, 4K: 7, 720p: 2) -- the usual choice of parallel
// encoders; the cost is well under 1 % of bitrate at these sizes.
VCP_HD int vcp_auto_slices(int mbh, int cabac) {
    if (!cabac) return 1;
    const int n = mbh / 17;
    return n < 1 ? 1 : n;
}

// transform size of an inter macroblock (High profile): cost4 / cost8 = sums of absolute 4x4 / 8x8
// Hadamard coefficients of the prediction residual; with orthonormal scaling the 8x8 sum weighs half
VCP_HD int vcp_prefer_8x8(int cost4, int cost8) { return cost8 < 2 * cost4; }

// macroblock types stored by the encoder
#define VCP_MB_I16 0
#define VCP_MB_P16 1
#define VCP_MB_PSKIP 2

// per-MB coefficient record, int16 units (levels are stored in zig-zag scan order)
#define VCP_LV_LUMA_DC 0        // 16: Intra16x16 DC levels
#define VCP_LV_LUMA 16          // 16 blocks (coding order) x 16
#define VCP_LV_CHROMA_DC 272    // 2 x 4
#define VCP_LV_CHROMA_AC 280    // 2 x 4 blocks x 16 (position 0 unused)
#define VCP_LV_STRIDE 408       // int16 per macroblock

// bit length of ue(k) / se(v): 2*floor(log2(codeNum+1)) + 1
VCP_HD int vcp_ue_len(unsigned k) {
    unsigned x = k + 1;
#if defined(__CUDA_ARCH__)
    return 2 * (31 - __clz((int)x)) + 1;
#else
    int n = 0;
    while (x > 1) { x >>= 1; n++; }
    return 2 * n + 1;
#endif
}
VCP_HD int vcp_se_len(int v) {
    return vcp_ue_len((v <= 0) ? (unsigned)(-2 * v) : (unsigned)(2 * v - 1));
}

// SAD-domain Lagrangian multiplier, ~ 2^((qp-12)/6)
VCP_HD int vcp_lambda(int qp) {
    // 64 * 2^(r/6) for r = 0..5: 64, 72, 81, 91, 102, 114 (one byte each, no local array)
    int q = qp < 12 ? 12 : qp;
    int e = (q - 12) / 6, r = (q - 12) % 6;
    int l = ((int)((0x72665B514840ull >> (8 * r)) & 255) << e) >> 6;
    return l < 1 ? 1 : l;
}

VCP_HD int vcp_median3(int a, int b, int c) {
    int mn = a < b ? a : b, mx = a < b ? b : a;
    return c < mn ? mn : (c > mx ? mx : c);
}
VCP_HD int vcp_clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
VCP_HD int vcp_clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
VCP_HD int vcp_iabs(int v) { return v < 0 ? -v : v; }

// ---- K1: colour conversion and scaling, integer only ------------------------------------------
// RGB -> YCbCr: BT.601 limited range (what swscale assumes for untagged RGB), 8-bit fixed point.
VCP_HD int vcp_rgb_y(int r, int g, int b) { return 16 + ((66 * r + 129 * g + 25 * b + 128) >> 8); }
VCP_HD int vcp_rgb_u(int r, int g, int b) { return 128 + ((-38 * r - 74 * g + 112 * b + 128) >> 8); }
VCP_HD int vcp_rgb_v(int r, int g, int b) { return 128 + ((112 * r - 94 * g - 18 * b + 128) >> 8); }
// Bilinear scaling, centre-aligned sample positions in 16.16 fixed point, 8-bit weights:
//   pos(d) = (2d+1) * in * 32768 / out - 32768, clamped to [0, (in-1)<<16]
VCP_HD long long vcp_scale_pos(int d, int in, int out) {
    long long p = ((long long)(2 * d + 1) * in * 32768) / out - 32768;
    const long long hi = (long long)(in - 1) << 16;
    return p < 0 ? 0 : (p > hi ? hi : p);
}
VCP_HD int vcp_bilerp(int p00, int p01, int p10, int p11, int fx, int fy) {
    return ((256 - fx) * (256 - fy) * p00 + fx * (256 - fy) * p01 + (256 - fx) * fy * p10 + fx * fy * p11 + 32768) >> 16;
}
// bytes of one input frame in format fmt (VCPENC_FMT_*: 0 yuv420p 1 nv12 2 rgb24 3 yuv444p 4 yuv422p 5 bgr24)
VCP_HD unsigned long long vcp_in_frame_bytes(int fmt, int w, int h) {
    const unsigned long long wh = (unsigned long long)w * h, c = (unsigned long long)((w + 1) / 2) * ((h + 1) / 2);
    switch (fmt) {
    case 0: case 1: return wh + 2 * c;
    case 2: case 5: case 3: return 3 * wh;
    case 4: return wh + 2ull * ((w + 1) / 2) * h;
    default: return 0;
    }
}

// CABAC is coded after the reconstruction chain (one sequential coder per slice, all pictures at
// once), so bitrate-targeted rate control is fed an estimate: bins * 10/16 bits (measured 0.60-0.66 bits per bin on the synthetic clips)
#define VCP_CABAC_BITS_PER_BIN_Q4 10

// ---- bitrate-targeted rate control (-b:v), integer only ------------------------------------
// GOPs are encoded independently and in parallel, so each GOP carries its own budget:
//   budget = bitrate * gop_frames / fps ; the IDR picture is expected to cost VCP_RC_I_WEIGHT
//   P pictures.  After picture t has been entropy coded, the QP of picture t+2 (the feedback
//   arrives two pictures late because entropy coding runs beside the reconstruction chain) is
//   qp0 + round(6*log2(bits spent / bits expected so far)), clamped.
#define VCP_RC_I_WEIGHT 6
#define VCP_RC_QP_I_OFFSET 3   // IDR pictures use qp - 3
#define VCP_RC_DOWN 14         // qp range relative to qp0
#define VCP_RC_UP 14
#define VCP_RC_STEP 2          // largest change from one picture to the next

// round(6*log2(num/den)), clamped to [-36, 36]
VCP_HD int vcp_rc_log2x6(unsigned long long num, unsigned long long den) {
    if (den == 0) den = 1;
    if (num == 0) return -36;
    unsigned long long r = (num << 16) / den;   // 16.16
    if (r == 0) return -36;
    int q = 0;
    while (r >= (2ull << 16) && q < 36) { r >>= 1; q += 6; }
    while (r < (1ull << 16) && q > -42) { r <<= 1; q -= 6; }
    const unsigned r1024 = (unsigned)(r >> 6);  // [1024, 2048)
    const unsigned mid[6] = {1085, 1218, 1367, 1534, 1722, 1933};  // 1024 * 2^((i+0.5)/6)
    for (int i = 0; i < 6; i++) q += r1024 >= mid[i];
    return q < -36 ? -36 : (q > 36 ? 36 : q);
}
// starting QP of every GOP from bits per pixel: 0.1 bpp <-> QP 28, -6 QP per doubling
VCP_HD int vcp_rc_initial_qp(int bitrate, int fps_num, int fps_den, int w, int h) {
    const unsigned long long bits_per_frame = (unsigned long long)bitrate * (unsigned)fps_den / (unsigned)(fps_num > 0 ? fps_num : 1);
    const int q = 28 - vcp_rc_log2x6(bits_per_frame * 10ull, (unsigned long long)w * (unsigned)h);
    return q < 14 ? 14 : (q > 45 ? 45 : q);
}
// QP of picture t+2, decided when the bits of picture t are known (the entropy coder runs beside
// the reconstruction chain, so feedback is two pictures late).  Rate model: bits ~ 2^(-QP/6).
//   target  = what is left of the GOP budget / P pictures left after t
//   bits_eq = bits of picture t as if it were a P picture (an IDR counts VCP_RC_I_WEIGHT pictures
//             and was coded VCP_RC_QP_I_OFFSET lower)
//   q       = qp(t) + 6 log2(bits_eq / target), moved at most VCP_RC_STEP away from qp(t+1),
//             kept within [qp0 - DOWN, qp0 + UP]
// Memoryless apart from qp(t), qp(t+1): no integrator to wind up behind the feedback delay; content
// whose rate/QP slope is flatter than the model converges geometrically from one side.
VCP_HD int vcp_rc_next_qp(int qp0, int qp_t, int qp_t1, int idr_t, unsigned long long bits_t,
                          unsigned long long cum_bits, int t, int L, unsigned long long gop_budget) {
    const int left = L - 1 - t;                      // pictures after t
    int q;
    if (left <= 0) q = qp_t1;
    else if (cum_bits >= gop_budget) q = qp_t1 + VCP_RC_STEP;
    else {
        const unsigned long long target = (gop_budget - cum_bits) / (unsigned)left;
        const unsigned long long bits_eq = idr_t ? bits_t / VCP_RC_I_WEIGHT : bits_t;
        const int base = idr_t ? qp_t + VCP_RC_QP_I_OFFSET : qp_t;
        q = base + vcp_rc_log2x6(bits_eq, target ? target : 1);
        q = (q + qp_t1 + 1) >> 1;                    // smooth: the answer to this choice arrives two pictures late
        q = q < qp_t1 - VCP_RC_STEP ? qp_t1 - VCP_RC_STEP : (q > qp_t1 + VCP_RC_STEP ? qp_t1 + VCP_RC_STEP : q);
    }
    q = q < qp0 - VCP_RC_DOWN ? qp0 - VCP_RC_DOWN : (q > qp0 + VCP_RC_UP ? qp0 + VCP_RC_UP : q);
    return q < 10 ? 10 : (q > 51 ? 51 : q);
}

// ---- VBV (-maxrate R -bufsize B), integer only ----------------------------------------------------
// Decoder buffer model: fullness F gains R/fps bits per picture interval, is capped at B, and loses the bits of a
// picture when it is decoded.  GOPs are encoded in parallel, so each GOP runs its own model from an ASSUMED start level
// of B/2 and is steered to end at least that full; F is monotone in its start value, so a stream whose first GOP really
// starts at B/2 or above (FFmpeg's default initial occupancy is 0.9 B) keeps every GOP at or above its model.
// Steering, two pictures ahead like the bitrate control: the QP chosen for picture t+2 is raised (by at most VCP_VBV_STEP
// over picture t+1) until the predicted fullness after picture t+2 stays above a floor that rises from B/10 at the GOP
// start to B/2 at its end.  The prediction prices pictures t+1 and t+2 from the bits of picture t with bits ~ 2^(-QP/6).
// An IDR picture that alone exceeds B/2 underflows the model (there is no second pass to re-encode it).
#define VCP_VBV_START_DIV 2
#define VCP_VBV_FLOOR_DIV 10
#define VCP_VBV_STEP 6
// bits * 2^(-dq/6): the same picture coded dq QP steps higher (lower when dq < 0)
VCP_HD unsigned long long vcp_rc_scale_bits(unsigned long long bits, int dq) {
    const unsigned f[6] = {65536, 58386, 52016, 46341, 41285, 36781};   // 65536 * 2^(-k/6)
    int sh = 0;
    if (dq < -48) dq = -48;
    if (dq > 48) dq = 48;
    while (dq < 0) { dq += 6; sh--; }
    while (dq >= 6) { dq -= 6; sh++; }
    if (bits > (1ull << 40)) bits = 1ull << 40;
    const unsigned long long v = (bits * f[dq]) >> 16;
    return sh >= 0 ? v >> sh : v << -sh;
}
VCP_HD long long vcp_vbv_advance(long long fullness, long long rate, long long buf, unsigned long long bits) {
    fullness += rate;
    if (fullness > buf) fullness = buf;
    return fullness - (long long)bits;
}
VCP_HD long long vcp_vbv_rate(int maxrate, int fps_num, int fps_den) {
    return (long long)((unsigned long long)maxrate * (unsigned)fps_den / (unsigned)(fps_num > 0 ? fps_num : 1));
}
// fullness: the model after picture t has been removed
VCP_HD int vcp_rc_vbv_qp(int q, int qp_t, int qp_t1, int idr_t, unsigned long long bits_t, long long fullness,
                         long long rate, long long buf, int t, int L) {
    const unsigned long long bits_eq = idr_t ? bits_t / VCP_RC_I_WEIGHT : bits_t;
    const int base = idr_t ? qp_t + VCP_RC_QP_I_OFFSET : qp_t;
    const long long f1 = vcp_vbv_advance(fullness, rate, buf, vcp_rc_scale_bits(bits_eq, qp_t1 - base));
    const long long lo = buf / VCP_VBV_FLOOR_DIV, hi = buf / VCP_VBV_START_DIV;
    const long long need = lo + (hi - lo) * (t + 3) / (L > 0 ? L : 1);
    const int qmax = qp_t1 + VCP_VBV_STEP > 51 ? 51 : qp_t1 + VCP_VBV_STEP;
    while (q < qmax && vcp_vbv_advance(f1, rate, buf, vcp_rc_scale_bits(bits_eq, q - base)) < need) q++;
    return q;
}
// the rate -b:v control aims for: -b:v, capped by -maxrate
VCP_HD int vcp_rc_eff_bitrate(int bitrate, int maxrate) { return maxrate > 0 && maxrate < bitrate ? maxrate : bitrate; }
// budget of a GOP of L pictures
VCP_HD unsigned long long vcp_rc_gop_budget(int bitrate, int maxrate, int fps_num, int fps_den, int L) {
    const int r = vcp_rc_eff_bitrate(bitrate, maxrate);
    return (unsigned long long)r * (unsigned)fps_den / (unsigned)(fps_num > 0 ? fps_num : 1) * (unsigned)L;
}
// One call per coded picture t of a GOP of L pictures, when its bits are known: returns the QP of picture t+2 and
// advances the GOP's state (*cum: bits spent, *fullness: VBV model).  abr: -b:v control (vcp_rc_next_qp) from qp0;
// otherwise constant QP qp_nom for P pictures, to which the QP returns VCP_RC_STEP per picture after a VBV excursion.
// vbv_buf = 0: no VBV.
VCP_HD int vcp_rc_picture(int abr, int qp0, int qp_nom, unsigned long long gop_budget, long long vbv_rate, long long vbv_buf,
                          int qp_t, int qp_t1, int idr_t, unsigned long long bits_t, int t, int L,
                          unsigned long long* cum, long long* fullness) {
    if (t == 0) { *cum = 0; *fullness = vbv_buf / VCP_VBV_START_DIV; }
    *cum += bits_t;
    int q;
    if (abr) q = vcp_rc_next_qp(qp0, qp_t, qp_t1, idr_t, bits_t, *cum, t, L, gop_budget);
    else q = qp_t1 - VCP_RC_STEP > qp_nom ? qp_t1 - VCP_RC_STEP : qp_nom;
    if (vbv_buf > 0) {
        *fullness = vcp_vbv_advance(*fullness, vbv_rate, vbv_buf, bits_t);
        q = vcp_rc_vbv_qp(q, qp_t, qp_t1, idr_t, bits_t, *fullness, vbv_rate, vbv_buf, t, L);
    }
    return q;
}
// does this parameter set run the per-picture feedback?  (VBV needs both -maxrate and -bufsize, like libx264)
VCP_HD int vcp_rc_has_vbv(int maxrate, int bufsize, int fps_num, int fps_den) { return maxrate > 0 && bufsize > 0 && fps_num > 0 && fps_den > 0; }

// ---- coefficient decimation of inter macroblocks (the rule x264 calls dct-decimate) ------------------------------
// Score of one transform block from the positions of its non-zero levels in scan order (`mask`, bit i = level i is
// non-zero): every level adds a weight that falls with the run of zeros in front of it -- an isolated +-1 after a long run
// is what the quantiser leaves of noise; a block with a level beyond +-1 is never decimated (score 9).  Scores add up
// over the four 4x4 blocks of an 8x8 luma group (an 8x8 transform block is its own group); a group scoring below 4 is
// zeroed, and so is the whole macroblock when the groups together score below 6.  Measured on the oracle (4x4 path):
// -2.6 % Bjontegaard rate on the hard clip, -0.7 % on the standard clip (profiles/r02_notes.md).
VCP_HD int vcp_decimate_score(unsigned long long mask, int big, int is8x8) {
    if (big) return 9;
    int sc = 0, prev = -1;
    for (int guard = 0; mask && sc < 9 && guard < 64; guard++) {
#ifdef __CUDA_ARCH__
        const int p = __ffsll((long long)mask) - 1;                              // lowest set bit
#else
        int p = 0;
        { unsigned long long m = mask; while (!(m & 1ull)) { m >>= 1; p++; } }
#endif
        const int run = p - prev - 1;
        sc += is8x8 ? (run < 4 ? 3 : run < 12 ? 2 : run < 32 ? 1 : 0) : (run < 1 ? 3 : run < 3 ? 2 : run < 6 ? 1 : 0);
        prev = p;
        mask &= mask - 1ull;
    }
    return sc > 9 ? 9 : sc;
}
VCP_HD int vcp_decimate_zero(int group_score, int mb_score) { return group_score < 4 || mb_score < 6; }

#endif  // VCP_ALGO_H
