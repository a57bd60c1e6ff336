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
    const int frac[6] = {64, 72, 81, 91, 102, 114};  // 64 * 2^(i/6)
    int q = qp < 12 ? 12 : qp;
    int e = (q - 12) / 6, r = (q - 12) % 6;
    int l = (frac[r] << e) >> 6;
    return l < 1 ? 1 : l;
}

VCP_HD int vcp_median3(int a, int b, int c) {
    int mn = a < b ? a : b, mx = a < b ? b : a;
    return c < mn ? mn : (c > mx ? mx : c);
}
VCP_HD int vcp_clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
VCP_HD int vcp_clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
VCP_HD int vcp_iabs(int v) { return v < 0 ? -v : v; }

#endif  // VCP_ALGO_H
