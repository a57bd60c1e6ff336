// Common device-side declarations: HBM layout, kernel launch prototypes.
//
// HBM layout (DESIGN.md "Data layout"):
//   planes are stored per frame with a replicated border so that motion vectors may leave
//   the picture; strides are multiples of 128 B and pixel (0,0) sits at a 16 B aligned
//   offset, so every macroblock row segment is a 16 B aligned vector.
//     luma   : stride ys, (ch + 2*VCP_PAD) rows, origin at row VCP_PAD, col VCP_PAD
//     chroma : stride cs = ys/2, (ch/2 + 2*VCP_PADC) rows, origin (VCP_PADC, VCP_PADC)
//     half   : stride hs = ys/2, (ch/2 + 2*VCP_PAD1) rows, origin (VCP_PAD1, VCP_PAD1)
//   per-macroblock records are structure-of-arrays, indexed [gop][mb].
#ifndef VCP_DEV_CUH
#define VCP_DEV_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "vcp_algo.h"

struct VcpGeom {
    int w, h;          // display size
    int cw, ch;        // coded size (multiples of 16)
    int mbw, mbh, nmb;
    int ys, cs, hs;    // strides in bytes
    int yoff, coff, hoff;          // byte offset of pixel (0,0) inside a plane
    size_t ysize, csize, hsize;    // bytes per plane
    int slices;
    int deblock_idc;
    int cabac;         // entropy_coding_mode_flag
    int t8x8;          // transform_8x8_mode_flag (High profile)
    int hevc;          // codec: 0 H.264, 1 HEVC (k6_hevc.cu; the motion search and the arithmetic coder are shared)
    int hevc_subpel;   // HEVC: half-sample luma motion from the 8-tap planes of k2_hpel.cu
    int hevc_sao;      // HEVC: sample adaptive offset (luma edge offsets)
    int effort;        // -preset tier: 0 fast (motion refine stops at half samples), 1 medium, 2 slow (= medium so far)
    // rate control (VCPENC_RC_ABR and / or the VBV model of -maxrate / -bufsize): see vcp_algo.h
    int rc_abr, rc_qp0, rc_bitrate, fps_num, fps_den;
    int rc_fb;                         // per-picture QP feedback on: rc_abr or VBV
    int rc_maxrate, rc_qp_nom;         // -maxrate; the P-picture QP of constant-QP mode
    long long rc_vbv_rate, rc_vbv_buf; // bits per picture interval, buffer size (0: no VBV)
};

__host__ __device__ __forceinline__ int vcp_slice_first_row(int s, int slices, int mbh) {
    return (int)(((long long)s * mbh) / slices);
}
__host__ __device__ __forceinline__ int vcp_slice_of_row(int row, int slices, int mbh) {
    int s = (int)(((long long)row * slices) / mbh);
    while (s + 1 < slices && vcp_slice_first_row(s + 1, slices, mbh) <= row) s++;
    while (s > 0 && vcp_slice_first_row(s, slices, mbh) > row) s--;
    return s;
}

// Frame addressing for one "step": frame index of GOP g at position t, recon ring slots.
struct VcpStep {
    int t;        // position inside the GOP
    int gop;      // GOP length
    int ngop;     // GOPs active in this step
    int ring;     // recon slots per GOP
    int nframes;  // total frames resident
    int gop0;     // index of the first resident GOP in the whole clip (idr_pic_id parity)
    int g0;       // first GOP of the group this launch covers (groups run on separate streams)
};
__host__ __device__ __forceinline__ int vcp_frame_of(const VcpStep& s, int g) { return g * s.gop + s.t; }
__host__ __device__ __forceinline__ int vcp_rec_slot(const VcpStep& s, int g, int t) { return g * s.ring + (t % s.ring); }
// luma planes per reconstruction slot: integer samples + the three half-sample planes (k2_hpel.cu)
#define VCP_REC_PLANES 4

// All device pointers of a session (plain struct passed by value to kernels).
struct VcpBufs {
    // originals, [nframes]
    uint8_t *src_y, *src_u, *src_v, *src_h;
    // reconstruction ring, [ngop_max * ring]
    uint8_t *rec_y, *rec_u, *rec_v;
    // pre-pass vectors, [nframes][nmb] (full-pel, x,y)
    short2* mvfp;
    // per-step macroblock records, [ngop_max][nmb]
    short2* mv;        // final quarter-pel vector
    short2* mvd;       // difference to the predictor
    uint8_t* mbtype;   // VCP_MB_*
    uint8_t* cbp;      // luma | chroma << 4
    uint8_t* modes;    // i16 mode | chroma mode << 2
    uint8_t* nnz;      // [..][24]: 16 luma raster, 4 Cb, 4 Cr
    int16_t* levels;   // [..][VCP_LV_STRIDE]
    uint32_t* mbbits;  // bits of the coded macroblock (incl. preceding skip run)
    uint32_t* mbbitoff;// bit offset inside the slice RBSP
    int32_t* skiprun;  // skipped macroblocks immediately preceding (same slice)
    // per frame
    uint8_t* qp;       // [nframes]
    // entropy output
    uint8_t* rbsp;         // [ngop_max][slices][rbsp_cap]  raw slice payloads of this step
    uint32_t* slice_bits;  // [ngop_max][slices] total RBSP bits (incl. header, trailing)
    uint8_t* out;          // NAL units, packed by an atomic cursor
    unsigned long long* out_cursor;
    uint2* out_index;      // [nframes][slices]: offset (low 32 of 64: see out_index_hi), size
    uint32_t* out_index_hi;
    uint32_t* frame_bits;  // [nframes] coded bits per frame (for rate control)
    int* error_flag;
    unsigned long long* rc_cum;  // [ngop_max] bits spent so far in the GOP
    long long* rc_full;          // [ngop_max] VBV model of the GOP
    int* db_sync;          // deblocking: [0] row ticket, [1 + gop*mbh + row] progress
    // CABAC (k5_cabac.cu): bins of every resident picture, per-macroblock descriptors, slice RBSPs
    uint16_t* bins;                  // arena, bump-allocated per macroblock
    unsigned long long* bins_cursor;
    size_t bins_cap;                 // in bins
    uint2* mbdesc;                   // [nframes][nmb]: offset low 32 | count (20 bits) + offset high << 20
    uint32_t* slice_bins;            // [nframes][slices] bins per slice
    uint16_t* sbins;                 // the same bins laid out as one contiguous, 16-byte aligned stream per slice (cabac_gather_kernel)
    unsigned long long* sbins_cursor;
    size_t sbins_cap;                // in bins
    unsigned long long* sslice_off;  // [nframes][slices] start of the slice's stream in sbins (bins)
    uint8_t* crbsp;                  // arena of slice RBSPs written by the arithmetic coder
    unsigned long long* crbsp_cursor;
    size_t crbsp_cap;
    unsigned long long* cslice_off;  // [nframes][slices]
    uint32_t* cslice_bytes;          // [nframes][slices]
    int* icount;                     // [ngop_max][slices] macroblocks of the step decided intra inside P pictures
    const uint32_t* rowinfo;  // [mbh]: first macroblock row of the row's slice | slice index << 16 (host-built)
    size_t rbsp_cap;
    size_t out_cap;
};

// first luma plane (integer samples G) of a reconstruction slot; B, H, J follow at +ysize each
__host__ __device__ __forceinline__ uint8_t* vcp_rec_luma(const VcpBufs& b, const VcpGeom& g, int slot) {
    return b.rec_y + (size_t)slot * VCP_REC_PLANES * g.ysize;
}

// ---- launchers (one per kernel family); all asynchronous on `st` -------------------------
void vcp_launch_k1_yuv420p(const uint8_t* in, size_t frame_bytes, int n0, int n, const VcpGeom& g,
                           const VcpBufs& b, cudaStream_t st);
void vcp_launch_k1_to_yuv420p(const uint8_t* in, size_t in_fb, int fmt, int w, int h, uint8_t* out, size_t out_fb, int n,
                              cudaStream_t st);
void vcp_launch_k1_scale(const uint8_t* in, size_t in_fb, int sw, int sh, uint8_t* out, size_t out_fb, int dw, int dh,
                         int n, cudaStream_t st);
struct VcpTmaps;   // vcp_tma.cuh: tensor maps over the picture planes (search windows are fetched by TMA)
void vcp_launch_me_prepass(const VcpGeom& g, const VcpBufs& b, const VcpTmaps& tm, int nframes, int gop, int t, cudaStream_t st, int g0 = 0, int g1 = -1);
void vcp_launch_me_refine(const VcpGeom& g, const VcpBufs& b, const VcpTmaps& tm, const VcpStep& s, cudaStream_t st);
void vcp_launch_p_recon(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_i_fix(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_i_recon(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_mbinfo(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_deblock(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_pad(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_hpel(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_cavlc_count(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_cavlc_scan(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_cavlc_write(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_nal_pack(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_cabac_bins(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_cabac_encode(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, int t0, int t1, cudaStream_t st);
void vcp_launch_hevc_p_recon(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_hevc_i_recon(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_hevc_i_fix(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_hevc_cuinfo(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_hevc_deblock(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_hevc_sao(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_hevc_sao_copy(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_hevc_bins(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);
void vcp_launch_rc_update(const VcpGeom& g, const VcpBufs& b, const VcpStep& s, cudaStream_t st);

#ifdef __CUDACC__
// slice geometry of a macroblock row without divisions (table built by the host at session create)
__device__ __forceinline__ int vcp_row_first(const VcpBufs& b, int my) { return (int)(__ldg(b.rowinfo + my) & 0xffff); }
__device__ __forceinline__ int vcp_row_slice(const VcpBufs& b, int my) { return (int)(__ldg(b.rowinfo + my) >> 16); }
// ---- small device helpers ------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_u32(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }
// 8 consecutive bytes starting at an arbitrary byte address (global or shared)
__device__ __forceinline__ uint2 ld8_unaligned(const uint8_t* p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
    uint2 r;
    r.x = __funnelshift_r(w0, w1, sh);
    r.y = __funnelshift_r(w1, w2, sh);
    return r;
}
__device__ __forceinline__ uint32_t ld4_unaligned(const uint8_t* p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    uint32_t sh = (uint32_t)(a & 3) * 8;
    return __funnelshift_r(q[0], q[1], sh);
}
// sum of absolute differences of the four bytes of a and b, plus c (one VABSDIFF4.U8.ACC)
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ uint32_t warp_min(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }
#endif

#endif  // VCP_DEV_CUH
