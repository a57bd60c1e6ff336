/* oracle/h264_oracle.c — CPU restatement of the hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's shared object.  It is never linked into libvcpenc and is not a
 * fallback of the product.
 *
 * What it restates.  The reference's hot path is `runFFmpegWithTimeout`
 * (/root/reference/cmd/consumer.go:370-394), which spawns a system `ffmpeg`; the codec
 * arithmetic is therefore in a third-party binary that is absent from /root/reference and
 * from this image (unpinned version, /root/reference/install.sh:90-99; encoders libx264 /
 * h264_nvenc selected by internal/config/config.go:45-50).  There is no reference source
 * file to follow and the reference holds no test or golden vector for this path:
 *
 *     PARITY UNPINNED against the reference's own encoder output.
 *
 * The oracle instead restates the *published* algorithm — ITU-T H.264 (intra 16x16 /
 * P_L0_16x16 / P_Skip prediction, 4x4 integer transform, quantisation, CAVLC, in-loop
 * deblocking, Annex-B) — as a scalar, sequential, macroblock-raster-order encoder, and is
 * pinned by the arbiter the north-star names: the FFmpeg h264 decoder (bundled libavcodec)
 * must decode its stream to exactly the reconstruction it reports (tests/test_oracle.py).
 * The CUDA path must then emit the same bytes as this oracle on the same input.
 *
 * Encoder-side (non-normative) decisions follow video_codec_pipeline_b200/csrc/vcp_algo.h.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VCP_TAB static const
#include "../video_codec_pipeline_b200/csrc/h264_tables.h"
#include "../video_codec_pipeline_b200/csrc/vcp_algo.h"
#include "../include/vcpenc.h"

/* ------------------------------------------------------------------------------------ */
/* bit writer                                                                            */
typedef struct {
    uint8_t* buf;
    size_t cap, pos; /* bytes */
    uint32_t cur;    /* bits accumulated */
    int nbits;
    int overflow;
} BW;

static void bw_init(BW* b, uint8_t* buf, size_t cap) {
    b->buf = buf; b->cap = cap; b->pos = 0; b->cur = 0; b->nbits = 0; b->overflow = 0;
}
static void bw_put(BW* b, int n, uint32_t v) { /* n <= 24 */
    while (n > 0) {
        int take = n > 8 ? 8 : n;
        uint32_t part = (v >> (n - take)) & ((1u << take) - 1);
        b->cur = (b->cur << take) | part;
        b->nbits += take;
        n -= take;
        while (b->nbits >= 8) {
            uint8_t byte = (uint8_t)(b->cur >> (b->nbits - 8));
            if (b->pos < b->cap) b->buf[b->pos++] = byte; else b->overflow = 1;
            b->nbits -= 8;
        }
    }
}
static void bw_put32(BW* b, uint32_t v) { bw_put(b, 16, v >> 16); bw_put(b, 16, v & 0xffff); }
static void bw_ue(BW* b, unsigned k) {
    unsigned x = k + 1; int n = 0;
    while ((x >> n) > 1) n++;
    bw_put(b, n, 0);
    bw_put(b, n + 1, x);
}
static void bw_se(BW* b, int v) { bw_ue(b, v <= 0 ? (unsigned)(-2 * v) : (unsigned)(2 * v - 1)); }
static void bw_trailing(BW* b) {
    bw_put(b, 1, 1);
    if (b->nbits) bw_put(b, 8 - b->nbits, 0);
}
static size_t bw_bits(const BW* b) { return b->pos * 8 + b->nbits; }

/* NAL unit: start code + header + emulation-prevented payload */
static size_t nal_write(uint8_t* out, size_t cap, int ref_idc, int type, const uint8_t* rbsp, size_t n) {
    size_t o = 0; int zeros = 0;
    if (cap < 5) return 0;
    out[o++] = 0; out[o++] = 0; out[o++] = 0; out[o++] = 1;
    out[o++] = (uint8_t)((ref_idc << 5) | type);
    for (size_t i = 0; i < n; i++) {
        if (zeros >= 2 && rbsp[i] <= 3) { if (o >= cap) return 0; out[o++] = 3; zeros = 0; }
        if (o >= cap) return 0;
        out[o++] = rbsp[i];
        zeros = rbsp[i] == 0 ? zeros + 1 : 0;
    }
    return o;
}

/* ------------------------------------------------------------------------------------ */
/* frames                                                                                */
typedef struct {
    int w, h;         /* coded luma size */
    int ys, cs;       /* strides */
    uint8_t *ybuf, *ubuf, *vbuf;
    uint8_t *y, *u, *v; /* pixel (0,0) */
} Frame;

static int frame_alloc(Frame* f, int w, int h) {
    f->w = w; f->h = h;
    f->ys = w + 2 * VCP_PAD; f->cs = w / 2 + 2 * VCP_PADC;
    f->ybuf = (uint8_t*)malloc((size_t)f->ys * (h + 2 * VCP_PAD));
    f->ubuf = (uint8_t*)malloc((size_t)f->cs * (h / 2 + 2 * VCP_PADC));
    f->vbuf = (uint8_t*)malloc((size_t)f->cs * (h / 2 + 2 * VCP_PADC));
    if (!f->ybuf || !f->ubuf || !f->vbuf) return -1;
    f->y = f->ybuf + (size_t)VCP_PAD * f->ys + VCP_PAD;
    f->u = f->ubuf + (size_t)VCP_PADC * f->cs + VCP_PADC;
    f->v = f->vbuf + (size_t)VCP_PADC * f->cs + VCP_PADC;
    return 0;
}
static void frame_free(Frame* f) { free(f->ybuf); free(f->ubuf); free(f->vbuf); }

static void plane_pad(uint8_t* p, int stride, int w, int h, int pad) {
    for (int y = 0; y < h; y++) {
        uint8_t* r = p + (size_t)y * stride;
        memset(r - pad, r[0], pad);
        memset(r + w, r[w - 1], pad);
    }
    for (int y = 1; y <= pad; y++) {
        memcpy(p - (size_t)y * stride - pad, p - pad, w + 2 * pad);
        memcpy(p + (size_t)(h - 1 + y) * stride - pad, p + (size_t)(h - 1) * stride - pad, w + 2 * pad);
    }
}
static void frame_pad(Frame* f) {
    plane_pad(f->y, f->ys, f->w, f->h, VCP_PAD);
    plane_pad(f->u, f->cs, f->w / 2, f->h / 2, VCP_PADC);
    plane_pad(f->v, f->cs, f->w / 2, f->h / 2, VCP_PADC);
}

/* K1 restated: load a display-size yuv420p picture into the coded-size padded frame
 * (replicating the last column/row up to the macroblock grid, then the border). */
static void plane_load(uint8_t* dst, int ds, int cw, int ch, const uint8_t* src, int sw, int sh) {
    for (int y = 0; y < ch; y++) {
        const uint8_t* s = src + (size_t)(y < sh ? y : sh - 1) * sw;
        uint8_t* d = dst + (size_t)y * ds;
        memcpy(d, s, sw);
        for (int x = sw; x < cw; x++) d[x] = s[sw - 1];
    }
}
static void frame_load_yuv420p(Frame* f, const uint8_t* src, int w, int h) {
    int cw = (w + 1) / 2, chh = (h + 1) / 2;
    plane_load(f->y, f->ys, f->w, f->h, src, w, h);
    plane_load(f->u, f->cs, f->w / 2, f->h / 2, src + (size_t)w * h, cw, chh);
    plane_load(f->v, f->cs, f->w / 2, f->h / 2, src + (size_t)w * h + (size_t)cw * chh, cw, chh);
    frame_pad(f);
}

/* K1 restated: any accepted input format -> tight yuv420p of the same size (vcp_algo.h) */
static void to_yuv420p(int fmt, const uint8_t* src, int w, int h, uint8_t* dst) {
    const int cw = (w + 1) / 2, ch = (h + 1) / 2;
    uint8_t *Y = dst, *U = dst + (size_t)w * h, *V = U + (size_t)cw * ch;
    if (fmt == VCPENC_FMT_YUV420P) { memcpy(dst, src, (size_t)w * h + 2 * (size_t)cw * ch); return; }
    if (fmt == VCPENC_FMT_NV12) {
        memcpy(Y, src, (size_t)w * h);
        const uint8_t* uv = src + (size_t)w * h;
        for (int i = 0; i < cw * ch; i++) { U[i] = uv[2 * i]; V[i] = uv[2 * i + 1]; }
        return;
    }
    if (fmt == VCPENC_FMT_RGB24 || fmt == VCPENC_FMT_BGR24) {
        const int ro = fmt == VCPENC_FMT_RGB24 ? 0 : 2, bo = 2 - ro;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const uint8_t* p = src + ((size_t)y * w + x) * 3;
                Y[(size_t)y * w + x] = (uint8_t)vcp_rgb_y(p[ro], p[1], p[bo]);
            }
        for (int y = 0; y < ch; y++)
            for (int x = 0; x < cw; x++) {
                int r = 0, g = 0, b = 0;
                for (int dy = 0; dy < 2; dy++)
                    for (int dx = 0; dx < 2; dx++) {
                        int yy = 2 * y + dy < h ? 2 * y + dy : h - 1, xx = 2 * x + dx < w ? 2 * x + dx : w - 1;
                        const uint8_t* p = src + ((size_t)yy * w + xx) * 3;
                        r += p[ro]; g += p[1]; b += p[bo];
                    }
                r = (r + 2) >> 2; g = (g + 2) >> 2; b = (b + 2) >> 2;
                U[(size_t)y * cw + x] = (uint8_t)vcp_rgb_u(r, g, b);
                V[(size_t)y * cw + x] = (uint8_t)vcp_rgb_v(r, g, b);
            }
        return;
    }
    memcpy(Y, src, (size_t)w * h);
    if (fmt == VCPENC_FMT_YUV444P) {
        for (int pl = 0; pl < 2; pl++) {
            const uint8_t* s = src + (size_t)w * h * (1 + pl);
            uint8_t* d = pl ? V : U;
            for (int y = 0; y < ch; y++)
                for (int x = 0; x < cw; x++) {
                    int y1 = 2 * y + 1 < h ? 2 * y + 1 : h - 1, x1 = 2 * x + 1 < w ? 2 * x + 1 : w - 1;
                    d[(size_t)y * cw + x] = (uint8_t)((s[(size_t)2 * y * w + 2 * x] + s[(size_t)2 * y * w + x1] +
                                                        s[(size_t)y1 * w + 2 * x] + s[(size_t)y1 * w + x1] + 2) >> 2);
                }
        }
    } else { /* yuv422p */
        for (int pl = 0; pl < 2; pl++) {
            const uint8_t* s = src + (size_t)w * h + (size_t)pl * cw * h;
            uint8_t* d = pl ? V : U;
            for (int y = 0; y < ch; y++)
                for (int x = 0; x < cw; x++) {
                    int y1 = 2 * y + 1 < h ? 2 * y + 1 : h - 1;
                    d[(size_t)y * cw + x] = (uint8_t)((s[(size_t)2 * y * cw + x] + s[(size_t)y1 * cw + x] + 1) >> 1);
                }
        }
    }
}
static void scale_plane(const uint8_t* s, int sw, int sh, uint8_t* d, int dw, int dh) {
    for (int y = 0; y < dh; y++) {
        long long py = vcp_scale_pos(y, sh, dh);
        int y0 = (int)(py >> 16), fy = (int)((py & 0xffff) >> 8), y1 = y0 + 1 < sh ? y0 + 1 : sh - 1;
        for (int x = 0; x < dw; x++) {
            long long px = vcp_scale_pos(x, sw, dw);
            int x0 = (int)(px >> 16), fx = (int)((px & 0xffff) >> 8), x1 = x0 + 1 < sw ? x0 + 1 : sw - 1;
            d[(size_t)y * dw + x] = (uint8_t)vcp_bilerp(s[(size_t)y0 * sw + x0], s[(size_t)y0 * sw + x1],
                                                        s[(size_t)y1 * sw + x0], s[(size_t)y1 * sw + x1], fx, fy);
        }
    }
}
static void scale_yuv420p(const uint8_t* s, int sw, int sh, uint8_t* d, int dw, int dh) {
    int scw = (sw + 1) / 2, sch = (sh + 1) / 2, dcw = (dw + 1) / 2, dch = (dh + 1) / 2;
    scale_plane(s, sw, sh, d, dw, dh);
    scale_plane(s + (size_t)sw * sh, scw, sch, d + (size_t)dw * dh, dcw, dch);
    scale_plane(s + (size_t)sw * sh + (size_t)scw * sch, scw, sch, d + (size_t)dw * dh + (size_t)dcw * dch, dcw, dch);
}

/* half-resolution luma with border VCP_PAD1, computed from the padded full-res plane */
typedef struct { int w, h, s; uint8_t* buf; uint8_t* p; } Half;
static int half_alloc(Half* hf, int w, int h) {
    hf->w = w / 2; hf->h = h / 2; hf->s = hf->w + 2 * VCP_PAD1;
    hf->buf = (uint8_t*)malloc((size_t)hf->s * (hf->h + 2 * VCP_PAD1));
    if (!hf->buf) return -1;
    hf->p = hf->buf + (size_t)VCP_PAD1 * hf->s + VCP_PAD1;
    return 0;
}
static void half_build(Half* hf, const Frame* f) {
    for (int y = -VCP_PAD1; y < hf->h + VCP_PAD1; y++)
        for (int x = -VCP_PAD1; x < hf->w + VCP_PAD1; x++) {
            const uint8_t* s = f->y + (size_t)(2 * y) * f->ys + 2 * x;
            hf->p[(size_t)y * hf->s + x] = (uint8_t)((s[0] + s[1] + s[f->ys] + s[f->ys + 1] + 2) >> 2);
        }
}

/* ------------------------------------------------------------------------------------ */
/* per-macroblock record                                                                  */
typedef struct {
    uint8_t type;        /* VCP_MB_* */
    uint8_t i16_mode, chroma_mode;
    uint8_t cbp;         /* luma bits 0..3 | chroma << 4 */
    int16_t mv[2];       /* quarter-pel */
    int16_t mvd[2];
    uint8_t nnz_y[16];   /* raster (y*4+x) total_coeff of each luma 4x4 (AC only for I16) */
    uint8_t nnz_c[2][4]; /* raster (y*2+x), AC */
    uint8_t t8x8;        /* transform_size_8x8_flag: luma coded as four 8x8 blocks.  Levels then sit in
                          * lv[VCP_LV_LUMA..] as the entropy coder wants them: CAVLC = the four interleaved
                          * 4x4 blocks of each 8x8 (block 4k+j, coefficient i <- scan position 4i+j, 9.2.1),
                          * CABAC = 64 scan positions per 8x8; nnz_y likewise (per 4x4 / whole 8x8) */
    int16_t lv[VCP_LV_STRIDE];
} MB;

/* ------------------------------------------------------------------------------------ */
/* transform / quant                                                                      */
static void fdct4(const int d[16], int w[16]) {
    int t[16];
    for (int i = 0; i < 4; i++) {
        int a0 = d[4 * i] + d[4 * i + 3], a1 = d[4 * i + 1] + d[4 * i + 2];
        int a2 = d[4 * i + 1] - d[4 * i + 2], a3 = d[4 * i] - d[4 * i + 3];
        t[4 * i] = a0 + a1; t[4 * i + 1] = 2 * a3 + a2; t[4 * i + 2] = a0 - a1; t[4 * i + 3] = a3 - 2 * a2;
    }
    for (int i = 0; i < 4; i++) {
        int a0 = t[i] + t[12 + i], a1 = t[4 + i] + t[8 + i];
        int a2 = t[4 + i] - t[8 + i], a3 = t[i] - t[12 + i];
        w[i] = a0 + a1; w[4 + i] = 2 * a3 + a2; w[8 + i] = a0 - a1; w[12 + i] = a3 - 2 * a2;
    }
}
/* inverse: input dequantised coefficients (raster), output residual (raster) */
static void idct4(const int c[16], int r[16]) {
    int t[16];
    for (int i = 0; i < 4; i++) {
        int e0 = c[4 * i] + c[4 * i + 2], e1 = c[4 * i] - c[4 * i + 2];
        int e2 = (c[4 * i + 1] >> 1) - c[4 * i + 3], e3 = c[4 * i + 1] + (c[4 * i + 3] >> 1);
        t[4 * i] = e0 + e3; t[4 * i + 1] = e1 + e2; t[4 * i + 2] = e1 - e2; t[4 * i + 3] = e0 - e3;
    }
    for (int i = 0; i < 4; i++) {
        int e0 = t[i] + t[8 + i], e1 = t[i] - t[8 + i];
        int e2 = (t[4 + i] >> 1) - t[12 + i], e3 = t[4 + i] + (t[12 + i] >> 1);
        r[i] = (e0 + e3 + 32) >> 6; r[4 + i] = (e1 + e2 + 32) >> 6;
        r[8 + i] = (e1 - e2 + 32) >> 6; r[12 + i] = (e0 - e3 + 32) >> 6;
    }
}
/* 8x8 forward transform (encoder's choice, matched by vcp_quant8_mf): rows, then columns */
static void fdct8_1d(const int x[8], int y[8]) {
    int a0 = x[0] + x[7], a1 = x[1] + x[6], a2 = x[2] + x[5], a3 = x[3] + x[4];
    int b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
    int a4 = x[0] - x[7], a5 = x[1] - x[6], a6 = x[2] - x[5], a7 = x[3] - x[4];
    int b4 = a5 + a6 + ((a4 >> 1) + a4), b5 = a4 - a7 - ((a6 >> 1) + a6);
    int b6 = a4 + a7 - ((a5 >> 1) + a5), b7 = a5 - a6 + ((a7 >> 1) + a7);
    y[0] = b0 + b1; y[1] = b4 + (b7 >> 2); y[2] = b2 + (b3 >> 1); y[3] = b5 + (b6 >> 2);
    y[4] = b0 - b1; y[5] = b6 - (b5 >> 2); y[6] = (b2 >> 1) - b3; y[7] = (b4 >> 2) - b7;
}
static void fdct8(const int d[64], int w[64]) {
    int t[64], in[8], out[8];
    for (int r = 0; r < 8; r++) { fdct8_1d(d + 8 * r, t + 8 * r); }
    for (int c = 0; c < 8; c++) {
        for (int r = 0; r < 8; r++) in[r] = t[8 * r + c];
        fdct8_1d(in, out);
        for (int r = 0; r < 8; r++) w[8 * r + c] = out[r];
    }
}
/* 8.5.13: one-dimensional inverse of the 8x8 transform */
static void idct8_1d(const int d[8], int o[8]) {
    int a0 = d[0] + d[4], a4 = d[0] - d[4], a2 = (d[2] >> 1) - d[6], a6 = d[2] + (d[6] >> 1);
    int b0 = a0 + a6, b2 = a4 + a2, b4 = a4 - a2, b6 = a0 - a6;
    int a1 = -d[3] + d[5] - d[7] - (d[7] >> 1), a3 = d[1] + d[7] - d[3] - (d[3] >> 1);
    int a5 = -d[1] + d[7] + d[5] + (d[5] >> 1), a7 = d[3] + d[5] + d[1] + (d[1] >> 1);
    int b1 = a1 + (a7 >> 2), b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5, b7 = a7 - (a1 >> 2);
    o[0] = b0 + b7; o[1] = b2 + b5; o[2] = b4 + b3; o[3] = b6 + b1;
    o[4] = b6 - b1; o[5] = b4 - b3; o[6] = b2 - b5; o[7] = b0 - b7;
}
static void idct8(const int c[64], int r[64]) {   /* each row first, then each column, then (x + 32) >> 6 */
    int t[64], in[8], out[8];
    for (int y = 0; y < 8; y++) idct8_1d(c + 8 * y, t + 8 * y);
    for (int x = 0; x < 8; x++) {
        for (int y = 0; y < 8; y++) in[y] = t[8 * y + x];
        idct8_1d(in, out);
        for (int y = 0; y < 8; y++) r[8 * y + x] = (out[y] + 32) >> 6;
    }
}
/* position class of normAdjust8x8 (8-318) for raster index i = y*8+x */
static int coef8_class(int i) {
    int y = i >> 3, x = i & 7;
    if (!(y & 3) && !(x & 3)) return 0;
    if ((y & 1) && (x & 1)) return 1;
    if ((y & 3) == 2 && (x & 3) == 2) return 2;
    if ((!(y & 3) && (x & 1)) || ((y & 1) && !(x & 3))) return 3;
    if ((!(y & 3) && (x & 3) == 2) || ((y & 3) == 2 && !(x & 3))) return 4;
    return 5;
}

static int quant1(int w, int mf, int f, int qbits) {
    int a = w < 0 ? -w : w;
    int l = (int)(((int64_t)a * mf + f) >> qbits);
    if (l > 2047) l = 2047;
    return w < 0 ? -l : l;
}
/* quantise 4x4 (raster w) into zig-zag levels; returns number of non-zeros among
 * scan positions [first..15] */
static int quant4x4(const int w[16], int qp, int intra, int first, int16_t lv[16]) {
    int qbits = 15 + qp / 6, f = (1 << qbits) / (intra ? 3 : 6), nz = 0;
    for (int k = first; k < 16; k++) {
        int i = vcp_zigzag4x4[k];
        int l = quant1(w[i], vcp_quant_mf[qp % 6][vcp_coef_class[i]], f, qbits);
        lv[k] = (int16_t)l;
        nz += l != 0;
    }
    return nz;
}
static void dequant4x4(const int16_t lv[16], int qp, int first, int c[16]) {
    for (int k = first; k < 16; k++) {
        int i = vcp_zigzag4x4[k];
        c[i] = (lv[k] * vcp_dequant_v[qp % 6][vcp_coef_class[i]]) << (qp / 6);
    }
}

/* ------------------------------------------------------------------------------------ */
/* inter prediction                                                                       */
/* 8x8: levels in zig-zag order; returns the non-zero count */
static int quant8x8(const int w[64], int qp, int intra, int16_t lv[64]) {
    int qbits = 16 + qp / 6, f = (1 << qbits) / (intra ? 3 : 6), nz = 0;
    for (int k = 0; k < 64; k++) {
        int i = vcp_zigzag8x8[k];
        lv[k] = (int16_t)quant1(w[i], vcp_quant8_mf[qp % 6][coef8_class(i)], f, qbits);
        nz += lv[k] != 0;
    }
    return nz;
}
/* 8.5.12.1 with flat scaling lists: LevelScale8x8 = 16 * normAdjust8x8 */
static void dequant8x8(const int16_t lv[64], int qp, int c[64]) {
    for (int k = 0; k < 64; k++) {
        int i = vcp_zigzag8x8[k];
        int ls = 16 * vcp_dequant8_v[qp % 6][coef8_class(i)];
        c[i] = qp >= 36 ? (lv[k] * ls) << (qp / 6 - 6) : (lv[k] * ls + (1 << (5 - qp / 6))) >> (6 - qp / 6);
    }
}

static inline int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }
static inline int hb1(const uint8_t* r, int s, int x, int y) { /* horizontal half between x,x+1 */
    const uint8_t* p = r + (size_t)y * s + x; (void)s;
    return tap6(p[-2], p[-1], p[0], p[1], p[2], p[3]);
}
static inline int hv1(const uint8_t* r, int s, int x, int y) { /* vertical half between y,y+1 */
    const uint8_t* p = r + (size_t)y * s + x;
    return tap6(p[-2 * s], p[-s], p[0], p[s], p[2 * s], p[3 * s]);
}
static inline int hj1(const uint8_t* r, int s, int x, int y) {
    return tap6(hb1(r, s, x, y - 2), hb1(r, s, x, y - 1), hb1(r, s, x, y), hb1(r, s, x, y + 1),
                hb1(r, s, x, y + 2), hb1(r, s, x, y + 3));
}
#define C5(v) vcp_clip255(((v) + 16) >> 5)
#define C10(v) vcp_clip255(((v) + 512) >> 10)
/* luma sample at quarter-pel position (8.4.2.2.1) */
static int luma_qpel(const uint8_t* r, int s, int ix, int iy, int fx, int fy) {
    int G = r[(size_t)iy * s + ix];
    switch (fy * 4 + fx) {
    case 0: return G;
    case 1: return (G + C5(hb1(r, s, ix, iy)) + 1) >> 1;
    case 2: return C5(hb1(r, s, ix, iy));
    case 3: return (r[(size_t)iy * s + ix + 1] + C5(hb1(r, s, ix, iy)) + 1) >> 1;
    case 4: return (G + C5(hv1(r, s, ix, iy)) + 1) >> 1;
    case 5: return (C5(hb1(r, s, ix, iy)) + C5(hv1(r, s, ix, iy)) + 1) >> 1;
    case 6: return (C5(hb1(r, s, ix, iy)) + C10(hj1(r, s, ix, iy)) + 1) >> 1;
    case 7: return (C5(hb1(r, s, ix, iy)) + C5(hv1(r, s, ix + 1, iy)) + 1) >> 1;
    case 8: return C5(hv1(r, s, ix, iy));
    case 9: return (C5(hv1(r, s, ix, iy)) + C10(hj1(r, s, ix, iy)) + 1) >> 1;
    case 10: return C10(hj1(r, s, ix, iy));
    case 11: return (C10(hj1(r, s, ix, iy)) + C5(hv1(r, s, ix + 1, iy)) + 1) >> 1;
    case 12: return (r[(size_t)(iy + 1) * s + ix] + C5(hv1(r, s, ix, iy)) + 1) >> 1;
    case 13: return (C5(hv1(r, s, ix, iy)) + C5(hb1(r, s, ix, iy + 1)) + 1) >> 1;
    case 14: return (C10(hj1(r, s, ix, iy)) + C5(hb1(r, s, ix, iy + 1)) + 1) >> 1;
    default: return (C5(hv1(r, s, ix + 1, iy)) + C5(hb1(r, s, ix, iy + 1)) + 1) >> 1;
    }
}
static void mc_luma16(const Frame* ref, int px, int py, int mvx, int mvy, uint8_t dst[256]) {
    int ix = px + (mvx >> 2), iy = py + (mvy >> 2), fx = mvx & 3, fy = mvy & 3;
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++)
            dst[y * 16 + x] = (uint8_t)luma_qpel(ref->y, ref->ys, ix + x, iy + y, fx, fy);
}
static void mc_chroma8(const uint8_t* r, int s, int px, int py, int mvx, int mvy, uint8_t dst[64]) {
    int ix = px + (mvx >> 3), iy = py + (mvy >> 3), dx = mvx & 7, dy = mvy & 7;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            const uint8_t* p = r + (size_t)(iy + y) * s + ix + x;
            dst[y * 8 + x] = (uint8_t)(((8 - dx) * (8 - dy) * p[0] + dx * (8 - dy) * p[1] +
                                        (8 - dx) * dy * p[s] + dx * dy * p[s + 1] + 32) >> 6);
        }
}
static int sad16(const uint8_t* a, int as, const uint8_t* b, int bs) {
    int s = 0;
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++) s += abs(a[y * as + x] - b[y * bs + x]);
    return s;
}

/* ------------------------------------------------------------------------------------ */
/* encoder context                                                                        */
typedef struct {
    vcpenc_params p;
    int mbw, mbh, nmb;
    int cw, ch; /* coded size */
    Frame cur, prev_orig, recon[2];
    Half hcur, hprev;
    MB* mbs;
    int16_t* mvfp; /* pre-pass full-pel mv per MB (x,y) */
    uint8_t* rbsp; size_t rbsp_cap;
} Enc;

static int slice_first_row(const Enc* e, int s) { return (int)((int64_t)s * e->mbh / e->p.slices); }
static int slice_of_row(const Enc* e, int row) {
    int s = (int)(((int64_t)row * e->p.slices) / e->mbh);
    while (s + 1 < e->p.slices && slice_first_row(e, s + 1) <= row) s++;
    while (s > 0 && slice_first_row(e, s) > row) s--;
    return s;
}

/* ---- K2a: motion search pre-pass on originals ---------------------------------------- */
static void me_prepass(Enc* e) {
    const Half *hc = &e->hcur, *hp = &e->hprev;
    for (int my = 0; my < e->mbh; my++)
        for (int mx = 0; mx < e->mbw; mx++) {
            /* L1 */
            uint32_t best = 0xffffffffu;
            const uint8_t* c = hc->p + (size_t)(8 * my) * hc->s + 8 * mx;
            int idx = 0;
            for (int dy = -VCP_ME_R1; dy <= VCP_ME_R1; dy++)
                for (int dx = -VCP_ME_R1; dx <= VCP_ME_R1; dx++, idx++) {
                    const uint8_t* r = hp->p + (size_t)(8 * my + dy) * hp->s + 8 * mx + dx;
                    int sad = 0;
                    for (int y = 0; y < 8; y++)
                        for (int x = 0; x < 8; x++) sad += abs(c[y * hc->s + x] - r[y * hp->s + x]);
                    uint32_t key = ((uint32_t)(sad + VCP_ME_L1_PEN * (abs(dx) + abs(dy))) << 16) | (uint32_t)idx;
                    if (key < best) best = key;
                }
            int bi = (int)(best & 0xffff), W = 2 * VCP_ME_R1 + 1;
            int cx = 2 * (bi % W - VCP_ME_R1), cy = 2 * (bi / W - VCP_ME_R1);
            /* L0 */
            best = 0xffffffffu; idx = 0;
            const uint8_t* c0 = e->cur.y + (size_t)(16 * my) * e->cur.ys + 16 * mx;
            for (int dy = -2; dy <= 2; dy++)
                for (int dx = -2; dx <= 2; dx++, idx++) {
                    int mvx = cx + dx, mvy = cy + dy;
                    const uint8_t* r = e->prev_orig.y + (size_t)(16 * my + mvy) * e->prev_orig.ys + 16 * mx + mvx;
                    int sad = sad16(c0, e->cur.ys, r, e->prev_orig.ys);
                    uint32_t key = ((uint32_t)(sad + VCP_ME_L0_PEN * (abs(mvx) + abs(mvy))) << 8) | (uint32_t)idx;
                    if (key < best) best = key;
                }
            bi = (int)(best & 0xff);
            e->mvfp[2 * (my * e->mbw + mx)] = (int16_t)(cx + bi % 5 - 2);
            e->mvfp[2 * (my * e->mbw + mx) + 1] = (int16_t)(cy + bi / 5 - 2);
        }
}

/* predictor estimate for the refine cost: median of the neighbours' PRE-PASS vectors */
static void pmv_estimate(const Enc* e, int mx, int my, int* px, int* py) {
    int row0 = slice_first_row(e, slice_of_row(e, my));
    int aA = mx > 0, aB = my > row0, aC = aB && mx + 1 < e->mbw, aD = aB && mx > 0;
    int ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0;
    const int16_t* m = e->mvfp;
    int i = my * e->mbw + mx;
    if (aA) { ax = 4 * m[2 * (i - 1)]; ay = 4 * m[2 * (i - 1) + 1]; }
    if (aB) { bx = 4 * m[2 * (i - e->mbw)]; by = 4 * m[2 * (i - e->mbw) + 1]; }
    if (aC) { cx = 4 * m[2 * (i - e->mbw + 1)]; cy = 4 * m[2 * (i - e->mbw + 1) + 1]; }
    else if (aD) { cx = 4 * m[2 * (i - e->mbw - 1)]; cy = 4 * m[2 * (i - e->mbw - 1) + 1]; }
    if (!aB && aA) { *px = ax; *py = ay; return; }
    *px = vcp_median3(ax, bx, cx); *py = vcp_median3(ay, by, cy);
}

/* ---- K2b: refine on the reconstructed reference ------------------------------------------ */
/* Intra-vs-inter decision for a P macroblock (vcp_algo.h: vcp_intra_wins).  The intra cost is an
 * ESTIMATE on the original picture (V / H / DC prediction of Intra16x16 from original neighbours),
 * so that it needs no reconstruction and every macroblock decides in parallel. */
static int intra_estimate(const Enc* e, int mx, int my) {
    int row0 = slice_first_row(e, slice_of_row(e, my));
    int aL = mx > 0, aT = my > row0;
    const uint8_t* c = e->cur.y + (size_t)(16 * my) * e->cur.ys + 16 * mx;
    const int s = e->cur.ys;
    int st = 0, sl = 0;
    for (int i = 0; i < 16; i++) { st += c[-s + i]; sl += c[i * s - 1]; }
    int dc = (aT && aL) ? (st + sl + 16) >> 5 : aT ? (st + 8) >> 4 : aL ? (sl + 8) >> 4 : 128;
    int sv = 0, sh = 0, sd = 0;
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++) {
            int p = c[y * s + x];
            sv += abs(p - c[-s + x]); sh += abs(p - c[y * s - 1]); sd += abs(p - dc);
        }
    int best = sd;
    if (aT && sv < best) best = sv;
    if (aL && sh < best) best = sh;
    return best;
}

static void me_refine_mb(Enc* e, const Frame* ref, int mx, int my, int qp, int16_t mv[2]) {
    int lam = vcp_lambda(qp), pmx, pmy;
    pmv_estimate(e, mx, my, &pmx, &pmy);
    int i = my * e->mbw + mx, px = 16 * mx, py = 16 * my;
    const uint8_t* c = e->cur.y + (size_t)py * e->cur.ys + px;
    int fx = e->mvfp[2 * i], fy = e->mvfp[2 * i + 1];
    /* full-pel candidates */
    int cand[11][2], n = 0;
    cand[n][0] = fx; cand[n][1] = fy; n++;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            if (!dx && !dy) continue;
            cand[n][0] = fx + dx; cand[n][1] = fy + dy; n++;
        }
    cand[n][0] = 0; cand[n][1] = 0; n++;
    cand[n][0] = vcp_clip3(-VCP_MV_FP_MAX, VCP_MV_FP_MAX, (pmx + 2) >> 2);
    cand[n][1] = vcp_clip3(-VCP_MV_FP_MAX, VCP_MV_FP_MAX, (pmy + 2) >> 2); n++;
    uint32_t best = 0xffffffffu;
    for (int k = 0; k < n; k++) {
        int vx = cand[k][0], vy = cand[k][1];
        const uint8_t* r = ref->y + (size_t)(py + vy) * ref->ys + px + vx;
        int cost = sad16(c, e->cur.ys, r, ref->ys) + lam * (vcp_se_len(4 * vx - pmx) + vcp_se_len(4 * vy - pmy));
        uint32_t key = ((uint32_t)cost << 4) | (uint32_t)k;
        if (key < best) best = key;
    }
    int bx = 4 * cand[best & 15][0], by = 4 * cand[best & 15][1];
    uint32_t bcost = best >> 4;
    e->mbs[i].type = VCP_MB_P16;
    if (bcost < VCP_SUBPEL_SKIP_COST) { mv[0] = (int16_t)bx; mv[1] = (int16_t)by; return; }
    /* half-pel then quarter-pel: 8 neighbours each, raster order, strict improvement.  -preset fast tiers (effort 0) stop
     * at half samples (vcpenc_params.effort: 0 fast, 1 medium, 2 slow; the built-in presets are medium / slow) */
    const int last_step = e->p.effort == 0 ? 2 : 1;
    for (int step = 2; step >= last_step; step--) {
        uint32_t sb = (bcost << 4) | 0; /* centre keeps priority (index 0) */
        int k = 1, cx = bx, cy = by;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                if (!dx && !dy) continue;
                int vx = cx + dx * step, vy = cy + dy * step;
                uint8_t pred[256];
                mc_luma16(ref, px, py, vx, vy, pred);
                int cost = sad16(c, e->cur.ys, pred, 16) + lam * (vcp_se_len(vx - pmx) + vcp_se_len(vy - pmy));
                uint32_t key = ((uint32_t)cost << 4) | (uint32_t)k;
                if (key < sb) { sb = key; bx = vx; by = vy; }
                k++;
            }
        bcost = sb >> 4;
    }
    mv[0] = (int16_t)bx; mv[1] = (int16_t)by;
    if (vcp_intra_wins(intra_estimate(e, mx, my), (int)bcost, lam)) e->mbs[i].type = VCP_MB_I16;
}

/* 4x4 or 8x8 transform for an inter macroblock?  Sum of absolute Hadamard coefficients of the
 * prediction residual, 4x4 blocks against 8x8 blocks (orthonormal scaling: |H8|/8 vs |H4|/4), the
 * transform that compacts the residual better wins (vcp_algo.h: vcp_prefer_8x8). */
static void hadamard4_2d(const int d[16], int h[16]) {
    int t[16];
    for (int i = 0; i < 4; i++) {
        int a = d[4 * i] + d[4 * i + 1], b = d[4 * i] - d[4 * i + 1], c = d[4 * i + 2] + d[4 * i + 3], e2 = d[4 * i + 2] - d[4 * i + 3];
        t[4 * i] = a + c; t[4 * i + 1] = b + e2; t[4 * i + 2] = a - c; t[4 * i + 3] = b - e2;
    }
    for (int i = 0; i < 4; i++) {
        int a = t[i] + t[4 + i], b = t[i] - t[4 + i], c = t[8 + i] + t[12 + i], e2 = t[8 + i] - t[12 + i];
        h[i] = a + c; h[4 + i] = b + e2; h[8 + i] = a - c; h[12 + i] = b - e2;
    }
}
static int prefer_8x8(const uint8_t* src, int ss, const uint8_t pred[256]) {
    long cost4 = 0, cost8 = 0;
    for (int k = 0; k < 4; k++) {
        int h[4][16];
        for (int j = 0; j < 4; j++) {
            int bx = (k & 1) * 8 + (j & 1) * 4, by = (k >> 1) * 8 + (j >> 1) * 4, d[16];
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++) d[4 * y + x] = src[(by + y) * ss + bx + x] - pred[(by + y) * 16 + bx + x];
            hadamard4_2d(d, h[j]);
            for (int i = 0; i < 16; i++) cost4 += abs(h[j][i]);
        }
        /* the 8x8 Hadamard of [[A,B],[C,D]] is the 4x4 Hadamard of A+-B+-C+-D */
        for (int i = 0; i < 16; i++) {
            int A = h[0][i], B = h[1][i], Cc = h[2][i], D = h[3][i];
            cost8 += abs(A + B + Cc + D) + abs(A - B + Cc - D) + abs(A + B - Cc - D) + abs(A - B - Cc + D);
        }
    }
    return vcp_prefer_8x8((int)cost4, (int)cost8);
}

/* ---- K3 (inter): predict, transform, quantise, reconstruct ------------------------------- */
static void chroma_dc_fwd_quant(int dc[4], int qpc, int intra, int16_t lv[4], int deq[4]) {
    /* 2x2 Hadamard, quantise (8.5.11.2 inverse restated forward), dequantise */
    int h[4] = {dc[0] + dc[1] + dc[2] + dc[3], dc[0] - dc[1] + dc[2] - dc[3],
                dc[0] + dc[1] - dc[2] - dc[3], dc[0] - dc[1] - dc[2] + dc[3]};
    int qbits = 15 + qpc / 6, f = (1 << qbits) / (intra ? 3 : 6);
    int mf = vcp_quant_mf[qpc % 6][0];
    for (int i = 0; i < 4; i++) lv[i] = (int16_t)quant1(h[i], mf, 2 * f, qbits + 1);
    int c[4] = {lv[0], lv[1], lv[2], lv[3]};
    int g[4] = {c[0] + c[1] + c[2] + c[3], c[0] - c[1] + c[2] - c[3],
                c[0] + c[1] - c[2] - c[3], c[0] - c[1] - c[2] + c[3]};
    int ls = 16 * vcp_dequant_v[qpc % 6][0];
    for (int i = 0; i < 4; i++) deq[i] = ((g[i] * ls) << (qpc / 6)) >> 5;
}

/* chroma of one MB (both planes): residual from pred, levels out, recon in place */
static void encode_chroma(Enc* e, MB* mb, Frame* rec, int mx, int my, int qp, int intra,
                          const uint8_t predu[64], const uint8_t predv[64]) {
    int qpc = vcp_chroma_qp[vcp_clip3(0, 51, qp)];
    int any_ac = 0, any_dc = 0;
    int coef[2][4][16];
    for (int pl = 0; pl < 2; pl++) {
        const uint8_t* src = (pl ? e->cur.v : e->cur.u) + (size_t)(8 * my) * e->cur.cs + 8 * mx;
        const uint8_t* pred = pl ? predv : predu;
        int dc[4];
        for (int b = 0; b < 4; b++) {
            int bx = (b & 1) * 4, by = (b >> 1) * 4, d[16], w[16];
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++)
                    d[y * 4 + x] = src[(by + y) * e->cur.cs + bx + x] - pred[(by + y) * 8 + bx + x];
            fdct4(d, w);
            dc[b] = w[0];
            int16_t* lv = mb->lv + VCP_LV_CHROMA_AC + (pl * 4 + b) * 16;
            lv[0] = 0;
            int nz = quant4x4(w, qpc, intra, 1, lv);
            mb->nnz_c[pl][b] = (uint8_t)nz;
            any_ac |= nz;
            coef[pl][b][0] = 0;
            dequant4x4(lv, qpc, 1, coef[pl][b]);
        }
        int deq[4];
        int16_t* dl = mb->lv + VCP_LV_CHROMA_DC + pl * 4;
        chroma_dc_fwd_quant(dc, qpc, intra, dl, deq);
        for (int b = 0; b < 4; b++) { coef[pl][b][0] = deq[b]; any_dc |= dl[b] != 0; }
    }
    int cbpc = any_ac ? 2 : (any_dc ? 1 : 0);
    mb->cbp = (uint8_t)((mb->cbp & 15) | (cbpc << 4));
    if (!any_ac) memset(mb->nnz_c, 0, sizeof mb->nnz_c);
    for (int pl = 0; pl < 2; pl++) {
        uint8_t* dst = (pl ? rec->v : rec->u) + (size_t)(8 * my) * rec->cs + 8 * mx;
        const uint8_t* pred = pl ? predv : predu;
        for (int b = 0; b < 4; b++) {
            int bx = (b & 1) * 4, by = (b >> 1) * 4, r[16];
            idct4(coef[pl][b], r);
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++)
                    dst[(by + y) * rec->cs + bx + x] = (uint8_t)vcp_clip255(pred[(by + y) * 8 + bx + x] + r[y * 4 + x]);
        }
    }
}

static void encode_p_mb(Enc* e, const Frame* ref, Frame* rec, int mx, int my, int qp) {
    MB* mb = &e->mbs[my * e->mbw + mx];
    if (mb->type == VCP_MB_I16) return;   /* decided intra by the refine: coded after the inter macroblocks */
    memset(mb->lv, 0, sizeof mb->lv);
    mb->type = VCP_MB_P16; mb->cbp = 0;
    uint8_t pred[256], pu[64], pv[64];
    mc_luma16(ref, 16 * mx, 16 * my, mb->mv[0], mb->mv[1], pred);
    mc_chroma8(ref->u, ref->cs, 8 * mx, 8 * my, mb->mv[0], mb->mv[1], pu);
    mc_chroma8(ref->v, ref->cs, 8 * mx, 8 * my, mb->mv[0], mb->mv[1], pv);
    const uint8_t* src = e->cur.y + (size_t)(16 * my) * e->cur.ys + 16 * mx;
    uint8_t* dst = rec->y + (size_t)(16 * my) * rec->ys + 16 * mx;
    mb->t8x8 = 0;
    /* Coefficient decimation (vcp_algo.h: vcp_decimate_*): an 8x8 luma group of an inter macroblock that holds only a
     * few isolated +-1 levels costs more bits than the distortion it removes -- such groups, or the whole macroblock,
     * are coded as zero (not in the fast -preset tiers, effort 0).  The scores walk the levels in scan order. */
    int score[4];
    if (e->p.transform8x8 && prefer_8x8(src, e->cur.ys, pred)) {
        /* High profile: the four 8x8 luma blocks of an inter macroblock */
        int16_t lv8[4][64];
        for (int k = 0; k < 4; k++) {
            int bx = (k & 1) * 8, by = (k >> 1) * 8, d[64], w[64];
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++)
                    d[y * 8 + x] = src[(by + y) * e->cur.ys + bx + x] - pred[(by + y) * 16 + bx + x];
            fdct8(d, w);
            quant8x8(w, qp, 0, lv8[k]);
            uint64_t mask = 0; int big = 0;
            for (int i = 0; i < 64; i++) { if (lv8[k][i]) mask |= 1ull << i; if (lv8[k][i] > 1 || lv8[k][i] < -1) big = 1; }
            score[k] = vcp_decimate_score(mask, big, 1);
        }
        const int tot = score[0] + score[1] + score[2] + score[3];
        for (int k = 0; k < 4; k++) {
            int bx = (k & 1) * 8, by = (k >> 1) * 8, c[64], r[64];
            if (e->p.effort > 0 && vcp_decimate_zero(score[k], tot)) memset(lv8[k], 0, sizeof lv8[k]);
            int nz = 0;
            for (int i = 0; i < 64; i++) nz += lv8[k][i] != 0;
            if (nz) mb->cbp |= (uint8_t)(1 << k);
            int cnt4[4] = {0, 0, 0, 0};
            for (int i = 0; i < 64; i++) {
                if (e->p.entropy) mb->lv[VCP_LV_LUMA + k * 64 + i] = lv8[k][i];
                else mb->lv[VCP_LV_LUMA + (k * 4 + (i & 3)) * 16 + (i >> 2)] = lv8[k][i];
                cnt4[i & 3] += lv8[k][i] != 0;
            }
            for (int j = 0; j < 4; j++)   /* luma4x4BlkIdx 4k+j */
                mb->nnz_y[vcp_blk_y[4 * k + j] * 4 + vcp_blk_x[4 * k + j]] = (uint8_t)(e->p.entropy ? nz : cnt4[j]);
            dequant8x8(lv8[k], qp, c);
            idct8(c, r);
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++)
                    dst[(by + y) * rec->ys + bx + x] = (uint8_t)vcp_clip255(pred[(by + y) * 16 + bx + x] + r[y * 8 + x]);
        }
        mb->t8x8 = (mb->cbp & 15) != 0;   /* the flag is only transmitted (else inferred 0) with coded luma */
    } else {
        score[0] = score[1] = score[2] = score[3] = 0;
        for (int b = 0; b < 16; b++) {
            int bx = vcp_blk_x[b] * 4, by = vcp_blk_y[b] * 4, d[16], w[16];
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++)
                    d[y * 4 + x] = src[(by + y) * e->cur.ys + bx + x] - pred[(by + y) * 16 + bx + x];
            fdct4(d, w);
            int16_t* lv = mb->lv + VCP_LV_LUMA + b * 16;
            quant4x4(w, qp, 0, 0, lv);
            uint64_t mask = 0; int big = 0;
            for (int i = 0; i < 16; i++) { if (lv[i]) mask |= 1ull << i; if (lv[i] > 1 || lv[i] < -1) big = 1; }
            score[b >> 2] += vcp_decimate_score(mask, big, 0);
        }
        const int tot = score[0] + score[1] + score[2] + score[3];
        for (int b = 0; b < 16; b++) {
            int bx = vcp_blk_x[b] * 4, by = vcp_blk_y[b] * 4, c[16], r[16];
            int16_t* lv = mb->lv + VCP_LV_LUMA + b * 16;
            if (e->p.effort > 0 && vcp_decimate_zero(score[b >> 2], tot)) memset(lv, 0, 16 * sizeof(int16_t));
            int nz = 0;
            for (int i = 0; i < 16; i++) nz += lv[i] != 0;
            mb->nnz_y[vcp_blk_y[b] * 4 + vcp_blk_x[b]] = (uint8_t)nz;
            if (nz) mb->cbp |= (uint8_t)(1 << (b >> 2));
            dequant4x4(lv, qp, 0, c);
            idct4(c, r);
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++)
                    dst[(by + y) * rec->ys + bx + x] = (uint8_t)vcp_clip255(pred[(by + y) * 16 + bx + x] + r[y * 4 + x]);
        }
    }
    encode_chroma(e, mb, rec, mx, my, qp, 0, pu, pv);
}

/* ---- motion vector prediction (8.4.1.3) and P_Skip (8.4.1.1) ------------------------------- */
static void mv_neighbours(const Enc* e, int mx, int my, int avail[4], int mv[4][2], int ref[4]) {
    /* 0:A left 1:B top 2:C top-right 3:D top-left; ref = 0 inter, -1 intra/unavailable */
    int row0 = slice_first_row(e, slice_of_row(e, my));
    int pos[4][2] = {{mx - 1, my}, {mx, my - 1}, {mx + 1, my - 1}, {mx - 1, my - 1}};
    for (int k = 0; k < 4; k++) {
        int x = pos[k][0], y = pos[k][1];
        avail[k] = x >= 0 && x < e->mbw && y >= row0 && y <= my;
        mv[k][0] = mv[k][1] = 0; ref[k] = -1;
        if (avail[k]) {
            const MB* n = &e->mbs[y * e->mbw + x];
            if (n->type != VCP_MB_I16) { ref[k] = 0; mv[k][0] = n->mv[0]; mv[k][1] = n->mv[1]; }
        }
    }
}
static void mvp16(const Enc* e, int mx, int my, int* px, int* py) {
    int av[4], mv[4][2], rf[4];
    mv_neighbours(e, mx, my, av, mv, rf);
    int c = av[2] ? 2 : 3;
    if (!av[1] && !av[c] && av[0]) { *px = mv[0][0]; *py = mv[0][1]; return; }
    int cnt = (rf[0] == 0) + (rf[1] == 0) + (rf[c] == 0);
    if (cnt == 1) {
        int k = rf[0] == 0 ? 0 : (rf[1] == 0 ? 1 : c);
        *px = mv[k][0]; *py = mv[k][1]; return;
    }
    *px = vcp_median3(mv[0][0], mv[1][0], mv[c][0]);
    *py = vcp_median3(mv[0][1], mv[1][1], mv[c][1]);
}
static void mv_pskip(const Enc* e, int mx, int my, int* px, int* py) {
    int av[4], mv[4][2], rf[4];
    mv_neighbours(e, mx, my, av, mv, rf);
    if (!av[0] || !av[1] || (rf[0] == 0 && !mv[0][0] && !mv[0][1]) || (rf[1] == 0 && !mv[1][0] && !mv[1][1])) {
        *px = *py = 0; return;
    }
    mvp16(e, mx, my, px, py);
}

/* ---- K3 (intra 16x16) -------------------------------------------------------------------- */
static void pred_i16(const Frame* rec, int mx, int my, int mode, int aL, int aT, uint8_t p[256]) {
    const uint8_t* o = rec->y + (size_t)(16 * my) * rec->ys + 16 * mx;
    int s = rec->ys;
    if (mode == 0) { for (int y = 0; y < 16; y++) memcpy(p + 16 * y, o - s, 16); }
    else if (mode == 1) { for (int y = 0; y < 16; y++) memset(p + 16 * y, o[y * s - 1], 16); }
    else if (mode == 2) {
        int sum = 0, dc;
        if (aT) for (int x = 0; x < 16; x++) sum += o[x - s];
        if (aL) for (int y = 0; y < 16; y++) sum += o[y * s - 1];
        dc = (aT && aL) ? (sum + 16) >> 5 : (aT || aL) ? (sum + 8) >> 4 : 128;
        memset(p, dc, 256);
    } else {
        int H = 0, V = 0;
        for (int i = 0; i < 8; i++) {
            H += (i + 1) * (o[8 + i - s] - o[6 - i - s]);
            V += (i + 1) * (o[(8 + i) * s - 1] - o[(6 - i) * s - 1]);
        }
        int a = 16 * (o[15 * s - 1] + o[15 - s]), b = (5 * H + 32) >> 6, c = (5 * V + 32) >> 6;
        for (int y = 0; y < 16; y++)
            for (int x = 0; x < 16; x++) p[y * 16 + x] = (uint8_t)vcp_clip255((a + b * (x - 7) + c * (y - 7) + 16) >> 5);
    }
}
static void pred_c8(const uint8_t* plane, int s, int mx, int my, int mode, int aL, int aT, uint8_t p[64]) {
    const uint8_t* o = plane + (size_t)(8 * my) * s + 8 * mx;
    if (mode == 0) { /* DC, per 4x4 */
        for (int b = 0; b < 4; b++) {
            int bx = (b & 1) * 4, by = (b >> 1) * 4, st = 0, sl = 0, dc;
            for (int i = 0; i < 4; i++) { if (aT) st += o[bx + i - s]; if (aL) sl += o[(by + i) * s - 1]; }
            if (b == 0 || b == 3) dc = (aT && aL) ? (st + sl + 4) >> 3 : aT ? (st + 2) >> 2 : aL ? (sl + 2) >> 2 : 128;
            else if (b == 1) dc = aT ? (st + 2) >> 2 : aL ? (sl + 2) >> 2 : 128;
            else dc = aL ? (sl + 2) >> 2 : aT ? (st + 2) >> 2 : 128;
            for (int y = 0; y < 4; y++) memset(p + (by + y) * 8 + bx, dc, 4);
        }
    } else if (mode == 1) { for (int y = 0; y < 8; y++) memset(p + 8 * y, o[y * s - 1], 8); }
    else if (mode == 2) { for (int y = 0; y < 8; y++) memcpy(p + 8 * y, o - s, 8); }
    else {
        int H = 0, V = 0;
        for (int i = 0; i < 4; i++) {
            H += (i + 1) * (o[4 + i - s] - o[2 - i - s]);
            V += (i + 1) * (o[(4 + i) * s - 1] - o[(2 - i) * s - 1]);
        }
        int a = 16 * (o[7 * s - 1] + o[7 - s]), b = (34 * H + 32) >> 6, c = (34 * V + 32) >> 6;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) p[y * 8 + x] = (uint8_t)vcp_clip255((a + b * (x - 3) + c * (y - 3) + 16) >> 5);
    }
}

static void encode_i_mb(Enc* e, Frame* rec, int mx, int my, int qp) {
    MB* mb = &e->mbs[my * e->mbw + mx];
    memset(mb->lv, 0, sizeof mb->lv);
    mb->type = VCP_MB_I16; mb->cbp = 0; mb->mv[0] = mb->mv[1] = 0; mb->mvd[0] = mb->mvd[1] = 0; mb->t8x8 = 0;
    int row0 = slice_first_row(e, slice_of_row(e, my));
    int aL = mx > 0, aT = my > row0;
    const uint8_t* src = e->cur.y + (size_t)(16 * my) * e->cur.ys + 16 * mx;
    /* luma mode: min SAD, ties to the lowest mode number */
    uint8_t pred[256], best_pred[256];
    uint32_t best = 0xffffffffu;
    for (int mode = 0; mode < 4; mode++) {
        if ((mode == 0 && !aT) || (mode == 1 && !aL) || (mode == 3 && !(aT && aL))) continue;
        pred_i16(rec, mx, my, mode, aL, aT, pred);
        uint32_t key = ((uint32_t)sad16(src, e->cur.ys, pred, 16) << 2) | (uint32_t)mode;
        if (key < best) { best = key; memcpy(best_pred, pred, 256); }
    }
    mb->i16_mode = (uint8_t)(best & 3);
    /* chroma mode: min SAD over both planes */
    uint8_t pu[64], pv[64], bpu[64], bpv[64];
    best = 0xffffffffu;
    for (int mode = 0; mode < 4; mode++) {
        if ((mode == 1 && !aL) || (mode == 2 && !aT) || (mode == 3 && !(aT && aL))) continue;
        pred_c8(rec->u, rec->cs, mx, my, mode, aL, aT, pu);
        pred_c8(rec->v, rec->cs, mx, my, mode, aL, aT, pv);
        int sad = 0;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) {
                sad += abs(e->cur.u[(size_t)(8 * my + y) * e->cur.cs + 8 * mx + x] - pu[y * 8 + x]);
                sad += abs(e->cur.v[(size_t)(8 * my + y) * e->cur.cs + 8 * mx + x] - pv[y * 8 + x]);
            }
        uint32_t key = ((uint32_t)sad << 2) | (uint32_t)mode;
        if (key < best) { best = key; memcpy(bpu, pu, 64); memcpy(bpv, pv, 64); }
    }
    mb->chroma_mode = (uint8_t)(best & 3);
    /* luma residual: 16 4x4 core transforms, DCs through a 4x4 Hadamard */
    int w[16][16], dcm[16]; /* dcm raster over 4x4 block grid */
    int any_ac = 0;
    for (int b = 0; b < 16; b++) {
        int bx = vcp_blk_x[b] * 4, by = vcp_blk_y[b] * 4, d[16];
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++)
                d[y * 4 + x] = src[(by + y) * e->cur.ys + bx + x] - best_pred[(by + y) * 16 + bx + x];
        fdct4(d, w[b]);
        dcm[vcp_blk_y[b] * 4 + vcp_blk_x[b]] = w[b][0];
        int16_t* lv = mb->lv + VCP_LV_LUMA + b * 16;
        lv[0] = 0;
        int nz = quant4x4(w[b], qp, 1, 1, lv);
        mb->nnz_y[vcp_blk_y[b] * 4 + vcp_blk_x[b]] = (uint8_t)nz;
        any_ac |= nz;
    }
    if (any_ac) mb->cbp = 15; else memset(mb->nnz_y, 0, 16);
    /* forward Hadamard of DCs (with /2), quantise */
    int t[16], hd[16];
    for (int i = 0; i < 4; i++) {
        int a0 = dcm[4 * i] + dcm[4 * i + 3], a1 = dcm[4 * i + 1] + dcm[4 * i + 2];
        int a2 = dcm[4 * i + 1] - dcm[4 * i + 2], a3 = dcm[4 * i] - dcm[4 * i + 3];
        t[4 * i] = a0 + a1; t[4 * i + 1] = a3 + a2; t[4 * i + 2] = a0 - a1; t[4 * i + 3] = a3 - a2;
    }
    for (int i = 0; i < 4; i++) {
        int a0 = t[i] + t[12 + i], a1 = t[4 + i] + t[8 + i];
        int a2 = t[4 + i] - t[8 + i], a3 = t[i] - t[12 + i];
        hd[i] = (a0 + a1) >> 1; hd[4 + i] = (a3 + a2) >> 1; hd[8 + i] = (a0 - a1) >> 1; hd[12 + i] = (a3 - a2) >> 1;
    }
    {
        int qbits = 15 + qp / 6, f = (1 << qbits) / 3, mf = vcp_quant_mf[qp % 6][0];
        int16_t* dl = mb->lv + VCP_LV_LUMA_DC;
        int c[16], g[16];
        for (int k = 0; k < 16; k++) {
            dl[k] = (int16_t)quant1(hd[vcp_zigzag4x4[k]], mf, 2 * f, qbits + 1);
            c[vcp_zigzag4x4[k]] = dl[k];
        }
        /* decoder side: inverse Hadamard then scale (8.5.10) */
        for (int i = 0; i < 4; i++) {
            int a0 = c[4 * i] + c[4 * i + 3], a1 = c[4 * i + 1] + c[4 * i + 2];
            int a2 = c[4 * i + 1] - c[4 * i + 2], a3 = c[4 * i] - c[4 * i + 3];
            t[4 * i] = a0 + a1; t[4 * i + 1] = a3 + a2; t[4 * i + 2] = a0 - a1; t[4 * i + 3] = a3 - a2;
        }
        for (int i = 0; i < 4; i++) {
            int a0 = t[i] + t[12 + i], a1 = t[4 + i] + t[8 + i];
            int a2 = t[4 + i] - t[8 + i], a3 = t[i] - t[12 + i];
            g[i] = a0 + a1; g[4 + i] = a3 + a2; g[8 + i] = a0 - a1; g[12 + i] = a3 - a2;
        }
        int ls = 16 * vcp_dequant_v[qp % 6][0];
        for (int i = 0; i < 16; i++)
            dcm[i] = qp >= 36 ? (g[i] * ls) << (qp / 6 - 6)
                              : (g[i] * ls + (1 << (5 - qp / 6))) >> (6 - qp / 6);
    }
    uint8_t* dst = rec->y + (size_t)(16 * my) * rec->ys + 16 * mx;
    for (int b = 0; b < 16; b++) {
        int bx = vcp_blk_x[b] * 4, by = vcp_blk_y[b] * 4, c[16], r[16];
        int16_t* lv = mb->lv + VCP_LV_LUMA + b * 16;
        if (!any_ac) memset(lv, 0, 32);
        dequant4x4(lv, qp, 1, c);
        c[0] = dcm[vcp_blk_y[b] * 4 + vcp_blk_x[b]];
        idct4(c, r);
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++)
                dst[(by + y) * rec->ys + bx + x] = (uint8_t)vcp_clip255(best_pred[(by + y) * 16 + bx + x] + r[y * 4 + x]);
    }
    encode_chroma(e, mb, rec, mx, my, qp, 1, bpu, bpv);
}

/* ------------------------------------------------------------------------------------ */
/* K4: deblocking (8.7), macroblock raster order                                           */
/* does the transform block holding 4x4 block `blk` (raster) carry coefficients? (8x8: the whole 8x8) */
static int blk_coded(const MB* m, int blk) {
    if (m->t8x8) return (m->cbp >> ((blk >> 3) * 2 + ((blk & 3) >> 1))) & 1;
    return m->nnz_y[blk] != 0;
}
static int bs_of(const MB* p, const MB* q, int pblk, int qblk, int mbedge) {
    if (p->type == VCP_MB_I16 || q->type == VCP_MB_I16) return mbedge ? 4 : 3;
    if (blk_coded(p, pblk) || blk_coded(q, qblk)) return 2;
    if (abs(p->mv[0] - q->mv[0]) >= 4 || abs(p->mv[1] - q->mv[1]) >= 4) return 1;
    return 0;
}
static void filter_luma_line(uint8_t* pix, int xs, int bS, int alpha, int beta, int tc0) {
    int p0 = pix[-xs], p1 = pix[-2 * xs], p2 = pix[-3 * xs], q0 = pix[0], q1 = pix[xs], q2 = pix[2 * xs];
    if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
    int ap = abs(p2 - p0), aq = abs(q2 - q0);
    if (bS < 4) {
        int tc = tc0 + (ap < beta) + (aq < beta);
        int d = vcp_clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
        pix[-xs] = (uint8_t)vcp_clip255(p0 + d);
        pix[0] = (uint8_t)vcp_clip255(q0 - d);
        if (ap < beta) pix[-2 * xs] = (uint8_t)(p1 + vcp_clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - 2 * p1) >> 1));
        if (aq < beta) pix[xs] = (uint8_t)(q1 + vcp_clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - 2 * q1) >> 1));
    } else {
        int p3 = pix[-4 * xs], q3 = pix[3 * xs];
        int small = abs(p0 - q0) < ((alpha >> 2) + 2);
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
static void filter_chroma_line(uint8_t* pix, int xs, int bS, int alpha, int beta, int tc0) {
    int p0 = pix[-xs], p1 = pix[-2 * xs], q0 = pix[0], q1 = pix[xs];
    if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
    if (bS < 4) {
        int tc = tc0 + 1;
        int d = vcp_clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
        pix[-xs] = (uint8_t)vcp_clip255(p0 + d);
        pix[0] = (uint8_t)vcp_clip255(q0 - d);
    } else {
        pix[-xs] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}
static void deblock_frame(Enc* e, Frame* f, int qp) {
    if (e->p.deblock_idc == 1) return;
    int qpc = vcp_chroma_qp[vcp_clip3(0, 51, qp)];
    int aY = vcp_alpha_tab[qp], bY = vcp_beta_tab[qp], aC = vcp_alpha_tab[qpc], bC = vcp_beta_tab[qpc];
    for (int my = 0; my < e->mbh; my++)
        for (int mx = 0; mx < e->mbw; mx++) {
            const MB* q = &e->mbs[my * e->mbw + mx];
            int row0 = slice_first_row(e, slice_of_row(e, my));
            int left_ok = mx > 0, top_ok = my > 0 && (e->p.deblock_idc != 2 || my > row0);
            uint8_t* Y = f->y + (size_t)(16 * my) * f->ys + 16 * mx;
            uint8_t* U = f->u + (size_t)(8 * my) * f->cs + 8 * mx;
            uint8_t* V = f->v + (size_t)(8 * my) * f->cs + 8 * mx;
            /* vertical edges (transform_size_8x8_flag: luma edges 1 and 3 are not transform edges) */
            for (int ed = 0; ed < 4; ed++) {
                if (ed == 0 && !left_ok) continue;
                if ((ed & 1) && q->t8x8) continue;
                const MB* p = ed == 0 ? q - 1 : q;
                int bS[4];
                for (int k = 0; k < 4; k++) bS[k] = bs_of(p, q, k * 4 + (ed == 0 ? 3 : ed - 1), k * 4 + ed, ed == 0);
                for (int y = 0; y < 16; y++) {
                    int b = bS[y >> 2];
                    if (b) filter_luma_line(Y + (size_t)y * f->ys + 4 * ed, 1, b, aY, bY, b < 4 ? vcp_tc0_tab[qp][b - 1] : 0);
                }
                if (!(ed & 1)) for (int y = 0; y < 8; y++) {
                    int b = bS[y >> 1];
                    if (!b) continue;
                    int t = b < 4 ? vcp_tc0_tab[qpc][b - 1] : 0;
                    filter_chroma_line(U + (size_t)y * f->cs + 2 * ed, 1, b, aC, bC, t);
                    filter_chroma_line(V + (size_t)y * f->cs + 2 * ed, 1, b, aC, bC, t);
                }
            }
            /* horizontal edges */
            for (int ed = 0; ed < 4; ed++) {
                if (ed == 0 && !top_ok) continue;
                if ((ed & 1) && q->t8x8) continue;
                const MB* p = ed == 0 ? q - e->mbw : q;
                int bS[4];
                for (int k = 0; k < 4; k++) bS[k] = bs_of(p, q, (ed == 0 ? 12 : 4 * (ed - 1)) + k, 4 * ed + k, ed == 0);
                for (int x = 0; x < 16; x++) {
                    int b = bS[x >> 2];
                    if (b) filter_luma_line(Y + (size_t)(4 * ed) * f->ys + x, f->ys, b, aY, bY, b < 4 ? vcp_tc0_tab[qp][b - 1] : 0);
                }
                if (!(ed & 1)) for (int x = 0; x < 8; x++) {
                    int b = bS[x >> 1];
                    if (!b) continue;
                    int t = b < 4 ? vcp_tc0_tab[qpc][b - 1] : 0;
                    filter_chroma_line(U + (size_t)(2 * ed) * f->cs + x, f->cs, b, aC, bC, t);
                    filter_chroma_line(V + (size_t)(2 * ed) * f->cs + x, f->cs, b, aC, bC, t);
                }
            }
        }
}

/* ------------------------------------------------------------------------------------ */
/* K5: CAVLC                                                                               */
/* residual_block_cavlc (7.3.5.3.2 / 9.2) for coefficients c[0..n-1] in scan order */
static int cavlc_block(BW* b, const int16_t* c, int n, int nC) {
    int total = 0, t1 = 0, last = -1;
    int lev[16], run[16];
    /* gather non-zero coefficients from the high-frequency end */
    int zeros = 0, total_zeros = 0;
    for (int i = n - 1; i >= 0; i--) {
        if (c[i]) {
            if (total) run[total - 1] = zeros;
            lev[total++] = c[i];
            zeros = 0;
            if (last < 0) last = i;
        } else if (total) zeros++;
    }
    if (total) { run[total - 1] = zeros; total_zeros = last + 1 - total; }
    for (int i = 0; i < total && i < 3; i++) { if (lev[i] == 1 || lev[i] == -1) t1++; else break; }
    /* coeff_token */
    if (nC == -1) bw_put(b, vcp_chroma_dc_coeff_token_len[4 * total + t1], vcp_chroma_dc_coeff_token_bits[4 * total + t1]);
    else {
        int tab = nC < 2 ? 0 : nC < 4 ? 1 : nC < 8 ? 2 : 3;
        bw_put(b, vcp_coeff_token_len[tab][4 * total + t1], vcp_coeff_token_bits[tab][4 * total + t1]);
    }
    if (!total) return 0;
    for (int i = 0; i < t1; i++) bw_put(b, 1, lev[i] < 0);
    int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
    for (int i = t1; i < total; i++) {
        int l = lev[i], code = l > 0 ? 2 * l - 2 : -2 * l - 1;
        if (i == t1 && t1 < 3) code -= 2;
        if (suffix_len == 0) {
            if (code < 14) bw_put(b, code + 1, 1);
            else if (code < 30) { bw_put(b, 15, 1); bw_put(b, 4, code - 14); }
            else { bw_put(b, 16, 1); bw_put(b, 12, code - 30); }
        } else {
            int pre = code >> suffix_len;
            if (pre < 15) { bw_put(b, pre + 1, 1); bw_put(b, suffix_len, code & ((1 << suffix_len) - 1)); }
            else { bw_put(b, 16, 1); bw_put(b, 12, code - (15 << suffix_len)); }
        }
        if (suffix_len == 0) suffix_len = 1;
        if (abs(l) > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
    }
    if (total < n) {
        if (n == 4) bw_put(b, vcp_chroma_dc_total_zeros_len[total - 1][total_zeros], vcp_chroma_dc_total_zeros_bits[total - 1][total_zeros]);
        else bw_put(b, vcp_total_zeros_len[total - 1][total_zeros], vcp_total_zeros_bits[total - 1][total_zeros]);
    }
    int left = total_zeros;
    for (int i = 0; i < total - 1 && left > 0; i++) {
        int zl = left > 7 ? 7 : left;
        bw_put(b, vcp_run_len[zl - 1][run[i]], vcp_run_bits[zl - 1][run[i]]);
        left -= run[i];
    }
    return total;
}

/* nC for luma block at raster (bx,by) of MB (mx,my) */
static int nnz_ctx(const Enc* e, int mx, int my, int bx, int by, int plane /*0 Y,1 Cb,2 Cr*/) {
    int row0 = slice_first_row(e, slice_of_row(e, my));
    int dim = plane ? 2 : 4;
    int nA = -1, nB = -1;
    const MB* m = &e->mbs[my * e->mbw + mx];
    if (bx > 0) nA = plane ? m->nnz_c[plane - 1][by * 2 + bx - 1] : m->nnz_y[by * 4 + bx - 1];
    else if (mx > 0) { const MB* l = m - 1; nA = plane ? l->nnz_c[plane - 1][by * 2 + 1] : l->nnz_y[by * 4 + 3]; }
    if (by > 0) nB = plane ? m->nnz_c[plane - 1][(by - 1) * 2 + bx] : m->nnz_y[(by - 1) * 4 + bx];
    else if (my > row0) { const MB* t = m - e->mbw; nB = plane ? t->nnz_c[plane - 1][(dim - 1) * 2 + bx] : t->nnz_y[12 + bx]; }
    if (nA >= 0 && nB >= 0) return (nA + nB + 1) >> 1;
    return nA >= 0 ? nA : nB >= 0 ? nB : 0;
}

static void cavlc_residual(BW* b, const Enc* e, const MB* mb, int mx, int my) {
    if (mb->type == VCP_MB_I16) cavlc_block(b, mb->lv + VCP_LV_LUMA_DC, 16, nnz_ctx(e, mx, my, 0, 0, 0));
    for (int blk = 0; blk < 16; blk++) {
        if (!(mb->cbp & (1 << (blk >> 2)))) continue;
        int nC = nnz_ctx(e, mx, my, vcp_blk_x[blk], vcp_blk_y[blk], 0);
        const int16_t* lv = mb->lv + VCP_LV_LUMA + blk * 16;
        if (mb->type == VCP_MB_I16) cavlc_block(b, lv + 1, 15, nC); else cavlc_block(b, lv, 16, nC);
    }
    int cc = mb->cbp >> 4;
    if (cc) for (int pl = 0; pl < 2; pl++) cavlc_block(b, mb->lv + VCP_LV_CHROMA_DC + pl * 4, 4, -1);
    if (cc & 2) for (int pl = 0; pl < 2; pl++)
        for (int blk = 0; blk < 4; blk++)
            cavlc_block(b, mb->lv + VCP_LV_CHROMA_AC + (pl * 4 + blk) * 16 + 1, 15, nnz_ctx(e, mx, my, blk & 1, blk >> 1, pl + 1));
}


/* ------------------------------------------------------------------------------------ */
/* K5 (CABAC): 9.3 — binarisation, context modelling and the arithmetic encoder, restated */
/* bit by bit from the standard (PutBit / RenormE / EncodeDecision / EncodeBypass /        */
/* EncodeFlush).  The CUDA coder uses a byte-wise equivalent; parity proves them equal.    */
#include "../video_codec_pipeline_b200/csrc/h264_cabac_tables.h"

typedef struct {
    BW* b;
    uint32_t low, range;
    int outstanding, first;
    uint8_t state[1024];   /* pStateIdx << 1 | valMPS */
    unsigned long long nbins;
} Cabac;

static void cabac_init(Cabac* c, BW* b, int tab /*0 I, 1+cabac_init_idc P*/, int qp) {
    c->b = b; c->low = 0; c->range = 510; c->outstanding = 0; c->first = 1; c->nbins = 0;
    for (int i = 0; i < 1024; i++) {
        int m = vcp_cabac_init_mn[tab][i][0], n = vcp_cabac_init_mn[tab][i][1];
        int pre = vcp_clip3(1, 126, ((m * vcp_clip3(0, 51, qp)) >> 4) + n);
        c->state[i] = pre <= 63 ? (uint8_t)((63 - pre) << 1) : (uint8_t)(((pre - 64) << 1) | 1);
    }
}
static void cabac_putbit(Cabac* c, int bit) {
    if (c->first) c->first = 0; else bw_put(c->b, 1, (uint32_t)bit);
    while (c->outstanding > 0) { bw_put(c->b, 1, (uint32_t)(1 - bit)); c->outstanding--; }
}
static void cabac_renorm(Cabac* c) {
    while (c->range < 256) {
        if (c->low < 256) cabac_putbit(c, 0);
        else if (c->low >= 512) { c->low -= 512; cabac_putbit(c, 1); }
        else { c->low -= 256; c->outstanding++; }
        c->range <<= 1; c->low <<= 1;
    }
}
static void cabac_encode(Cabac* c, int ctx, int bin) {
    int st = c->state[ctx] >> 1, mps = c->state[ctx] & 1;
    uint32_t rlps = vcp_cabac_range_lps[st][(c->range >> 6) & 3];
    c->range -= rlps;
    if (bin != mps) {
        c->low += c->range; c->range = rlps;
        if (st == 0) mps ^= 1;
        st = vcp_cabac_trans_lps[st];
    } else if (st < 62) st++;
    c->state[ctx] = (uint8_t)((st << 1) | mps);
    cabac_renorm(c);
    c->nbins++;
}
static void cabac_bypass(Cabac* c, int bin) {
    c->low <<= 1;
    if (bin) c->low += c->range;
    if (c->low >= 1024) { cabac_putbit(c, 1); c->low -= 1024; }
    else if (c->low < 512) cabac_putbit(c, 0);
    else { c->low -= 512; c->outstanding++; }
    c->nbins++;
}
static void cabac_terminate(Cabac* c, int bin) {
    c->range -= 2;
    if (bin) {
        c->low += c->range;
        c->range = 2;
        cabac_renorm(c);
        cabac_putbit(c, (c->low >> 9) & 1);
        bw_put(c->b, 2, ((c->low >> 7) & 3) | 1);   /* the final 1 is rbsp_stop_one_bit */
    } else cabac_renorm(c);
    c->nbins++;
}

/* unsigned Exp-Golomb of order k in bypass bins (9.3.2.3 suffix) */
static void cabac_ueg_bypass(Cabac* c, unsigned v, int k) {
    while (v >= (1u << k)) { cabac_bypass(c, 1); v -= 1u << k; k++; }
    cabac_bypass(c, 0);
    while (k--) cabac_bypass(c, (int)((v >> k) & 1));
}

/* flags a neighbour contributes to context selection */
static int mb_dc_cbf(const MB* m, int which /*0 luma DC, 1 Cb DC, 2 Cr DC*/) {
    if (which == 0) {
        if (m->type != VCP_MB_I16) return 0;
        for (int i = 0; i < 16; i++) if (m->lv[VCP_LV_LUMA_DC + i]) return 1;
        return 0;
    }
    if (m->type == VCP_MB_PSKIP || !(m->cbp >> 4)) return 0;
    for (int i = 0; i < 4; i++) if (m->lv[VCP_LV_CHROMA_DC + (which - 1) * 4 + i]) return 1;
    return 0;
}

/* residual_block_cabac (7.3.5.3.3) for coefficients c[0..n-1] in scan order */
static void cabac_block(Cabac* cb, const int16_t* c, int n, int cat, int cbf_inc) {
    /* ctxBlockCat 0..4 (tables 9-34, 9-40); 5 = luma 8x8: own context ranges 402 / 417 / 426, position
     * dependent increments (table 9-43), no coded_block_flag (inferred from the coded block pattern) */
    static const int cbf_off[5] = {0, 4, 8, 12, 16}, sig_off[5] = {0, 15, 29, 44, 47}, abs_off[6] = {0, 10, 20, 30, 39, 199};
    int last = -1;
    for (int i = 0; i < n; i++) if (c[i]) last = i;
    if (cat != 5) cabac_encode(cb, 85 + cbf_off[cat] + cbf_inc, last >= 0);
    if (last < 0) return;
    for (int i = 0; i < n - 1; i++) {
        int sctx, lctx;
        if (cat == 5) { sctx = 402 + vcp_cabac_sig8x8[i]; lctx = 417 + vcp_cabac_last8x8[i]; }
        else { int inc = cat == 3 ? (i < 2 ? i : 2) : i; sctx = 105 + sig_off[cat] + inc; lctx = 166 + sig_off[cat] + inc; }
        cabac_encode(cb, sctx, c[i] != 0);
        if (c[i]) {
            cabac_encode(cb, lctx, i == last);
            if (i == last) break;
        }
    }
    int gt1 = 0, eq1 = 0;
    for (int i = last; i >= 0; i--) {
        if (!c[i]) continue;
        int a = abs(c[i]) - 1;   /* coeff_abs_level_minus1: prefix TU cMax 14, suffix EG0 */
        int inc = gt1 ? 0 : (1 + eq1 < 4 ? 1 + eq1 : 4);
        cabac_encode(cb, 227 + abs_off[cat] + inc, a > 0);
        if (a > 0) {
            int lim = 4 - (cat == 3);
            int ctx = 227 + abs_off[cat] + 5 + (gt1 < lim ? gt1 : lim);
            for (int k = 1; k < (a < 14 ? a : 14); k++) cabac_encode(cb, ctx, 1);
            if (a < 14) cabac_encode(cb, ctx, 0);
            else cabac_ueg_bypass(cb, (unsigned)(a - 14), 0);
            gt1++;
        } else eq1++;
        cabac_bypass(cb, c[i] < 0);
    }
}

static void cabac_mvd(Cabac* c, int base, int v, int amvd) {
    int a = abs(v);
    cabac_encode(c, base + (amvd < 3 ? 0 : amvd > 32 ? 2 : 1), a > 0);
    if (!a) return;
    /* prefix TU cMax 9: bins 1.. use ctx base+3,+4,+5,+6,+6,... ; suffix EG3 ; sign */
    for (int k = 1; k < (a < 9 ? a : 9); k++) cabac_encode(c, base + 3 + (k - 1 < 3 ? k - 1 : 3), 1);
    if (a < 9) cabac_encode(c, base + 3 + (a - 1 < 3 ? a - 1 : 3), 0);
    else cabac_ueg_bypass(c, (unsigned)(a - 9), 3);
    cabac_bypass(c, v < 0);
}

static void cabac_i16_type(Cabac* c, const MB* mb, int ctx0, int base, int islice) {
    /* prefix bin "not I_NxN" at ctx0, then terminate(0) = not I_PCM, then 5 fields (9.3.2.5) */
    cabac_encode(c, ctx0, 1);
    cabac_terminate(c, 0);
    int cc = mb->cbp >> 4;
    cabac_encode(c, base + 1, (mb->cbp & 15) != 0);
    cabac_encode(c, base + 2, cc != 0);
    if (cc) cabac_encode(c, base + 2 + islice, cc == 2);
    cabac_encode(c, base + 3 + islice, mb->i16_mode >> 1);
    cabac_encode(c, base + 3 + 2 * islice, mb->i16_mode & 1);
}

/* slice data (7.3.4) with entropy_coding_mode_flag = 1 for MB rows [r0,r1) */
static void write_slice_data_cabac(Enc* e, BW* b, int r0, int r1, int idr, int qp, unsigned long long* nbins) {
    while (b->nbits) bw_put(b, 1, 1);   /* cabac_alignment_one_bit */
    Cabac c; cabac_init(&c, b, idr ? 0 : 1, qp);
    for (int my = r0; my < r1; my++)
        for (int mx = 0; mx < e->mbw; mx++) {
            const MB* mb = &e->mbs[my * e->mbw + mx];
            const MB* A = mx > 0 ? mb - 1 : NULL;
            const MB* B = my > r0 ? mb - e->mbw : NULL;
            const int intra = mb->type == VCP_MB_I16;
            if (!idr) {
                int inc = (A && A->type != VCP_MB_PSKIP) + (B && B->type != VCP_MB_PSKIP);
                cabac_encode(&c, 11 + inc, mb->type == VCP_MB_PSKIP);
            }
            if (mb->type != VCP_MB_PSKIP) {
                if (idr) {
                    /* neighbours here are always Intra16x16: condTermFlag = available */
                    cabac_i16_type(&c, mb, 3 + (A != NULL) + (B != NULL), 5, 1);
                } else if (intra) {
                    cabac_encode(&c, 14, 1);
                    cabac_i16_type(&c, mb, 17, 17, 0);
                } else {
                    cabac_encode(&c, 14, 0); cabac_encode(&c, 15, 0); cabac_encode(&c, 16, 0);   /* P_L0_16x16 */
                }
                if (intra) {
                    int inc = (A && A->type == VCP_MB_I16 && A->chroma_mode) + (B && B->type == VCP_MB_I16 && B->chroma_mode);
                    cabac_encode(&c, 64 + inc, mb->chroma_mode > 0);
                    if (mb->chroma_mode > 0) {
                        cabac_encode(&c, 67, mb->chroma_mode > 1);
                        if (mb->chroma_mode > 1) cabac_encode(&c, 67, mb->chroma_mode > 2);
                    }
                } else {
                    for (int k = 0; k < 2; k++) {
                        int am = (A && A->type == VCP_MB_P16 ? abs(A->mvd[k]) : 0) + (B && B->type == VCP_MB_P16 ? abs(B->mvd[k]) : 0);
                        cabac_mvd(&c, k ? 47 : 40, mb->mvd[k], am);
                    }
                    /* coded_block_pattern: 4 luma bins (8x8 raster), then chroma */
                    int cbpA = A ? (A->type == VCP_MB_PSKIP ? 0 : A->cbp) : 0x0f, cbpB = B ? (B->type == VCP_MB_PSKIP ? 0 : B->cbp) : 0x0f;
                    int cur = mb->cbp;
                    for (int k = 0; k < 4; k++) {
                        int a = (k & 1) ? (cur >> (k - 1)) & 1 : (cbpA >> (k + 1)) & 1;
                        int bb = (k & 2) ? (cur >> (k - 2)) & 1 : (cbpB >> (k + 2)) & 1;
                        cabac_encode(&c, 73 + !a + 2 * !bb, (cur >> k) & 1);
                    }
                    int ca = A ? (A->type == VCP_MB_PSKIP ? 0 : A->cbp >> 4) : 0, cbb = B ? (B->type == VCP_MB_PSKIP ? 0 : B->cbp >> 4) : 0;
                    cabac_encode(&c, 77 + (ca > 0) + 2 * (cbb > 0), (cur >> 4) > 0);
                    if (cur >> 4) cabac_encode(&c, 77 + 4 + (ca == 2) + 2 * (cbb == 2), (cur >> 4) == 2);
                }
                if (!intra && e->p.transform8x8 && (mb->cbp & 15))
                    cabac_encode(&c, 399 + (A && A->t8x8) + (B && B->t8x8), mb->t8x8);   /* transform_size_8x8_flag */
                if (intra || mb->cbp) cabac_encode(&c, 60, 0);   /* mb_qp_delta = 0 (and so was the previous one) */
                /* residual */
                const int un = intra ? 1 : 0;   /* flag assumed for an unavailable neighbour */
                if (intra)
                    cabac_block(&c, mb->lv + VCP_LV_LUMA_DC, 16, 0, (A ? mb_dc_cbf(A, 0) : un) + 2 * (B ? mb_dc_cbf(B, 0) : un));
                if (mb->t8x8) {
                    for (int k = 0; k < 4; k++)
                        if (mb->cbp & (1 << k)) cabac_block(&c, mb->lv + VCP_LV_LUMA + k * 64, 64, 5, 0);
                } else
                for (int blk = 0; blk < 16; blk++) {
                    if (!(mb->cbp & (1 << (blk >> 2)))) continue;
                    int bx = vcp_blk_x[blk], by = vcp_blk_y[blk];
                    int fa = bx > 0 ? mb->nnz_y[by * 4 + bx - 1] != 0 : A ? (A->type != VCP_MB_PSKIP && A->nnz_y[by * 4 + 3] != 0) : un;
                    int fb = by > 0 ? mb->nnz_y[(by - 1) * 4 + bx] != 0 : B ? (B->type != VCP_MB_PSKIP && B->nnz_y[12 + bx] != 0) : un;
                    const int16_t* lv = mb->lv + VCP_LV_LUMA + blk * 16;
                    if (intra) cabac_block(&c, lv + 1, 15, 1, fa + 2 * fb); else cabac_block(&c, lv, 16, 2, fa + 2 * fb);
                }
                int cc = mb->cbp >> 4;
                if (cc) for (int pl = 0; pl < 2; pl++)
                    cabac_block(&c, mb->lv + VCP_LV_CHROMA_DC + pl * 4, 4, 3,
                                (A ? mb_dc_cbf(A, 1 + pl) : un) + 2 * (B ? mb_dc_cbf(B, 1 + pl) : un));
                if (cc & 2) for (int pl = 0; pl < 2; pl++)
                    for (int blk = 0; blk < 4; blk++) {
                        int bx = blk & 1, by = blk >> 1;
                        int fa = bx > 0 ? mb->nnz_c[pl][by * 2] != 0 : A ? (A->type != VCP_MB_PSKIP && A->nnz_c[pl][by * 2 + 1] != 0) : un;
                        int fb = by > 0 ? mb->nnz_c[pl][bx] != 0 : B ? (B->type != VCP_MB_PSKIP && B->nnz_c[pl][2 + bx] != 0) : un;
                        cabac_block(&c, mb->lv + VCP_LV_CHROMA_AC + (pl * 4 + blk) * 16 + 1, 15, 4, fa + 2 * fb);
                    }
            }
            cabac_terminate(&c, my == r1 - 1 && mx == e->mbw - 1);   /* end_of_slice_flag */
        }
    while (b->nbits) bw_put(b, 1, 0);   /* rbsp_alignment_zero_bit (stop bit came from the flush) */
    if (nbins) *nbins += c.nbins;
}

/* ------------------------------------------------------------------------------------ */
/* headers                                                                                */
static int level_idc_for(int mbw, int mbh, int fps_num, int fps_den) {
    /* Table A-1: MaxMBPS, MaxFS */
    static const struct { int idc; long mbps; long fs; } L[] = {
        {10, 1485, 99}, {11, 3000, 396}, {12, 6000, 396}, {13, 11880, 396}, {20, 11880, 396},
        {21, 19800, 792}, {22, 20250, 1620}, {30, 40500, 1620}, {31, 108000, 3600},
        {32, 216000, 5120}, {40, 245760, 8192}, {42, 522240, 8704}, {50, 589824, 22080},
        {51, 983040, 36864}, {52, 2073600, 36864}};
    long fs = (long)mbw * mbh;
    long mbps = (long)((double)fs * fps_num / (fps_den > 0 ? fps_den : 1) + 0.5);
    for (unsigned i = 0; i < sizeof L / sizeof L[0]; i++)
        if (fs <= L[i].fs && mbps <= L[i].mbps) return L[i].idc;
    return 52;
}
static size_t write_sps(const Enc* e, uint8_t* out, size_t cap) {
    uint8_t tmp[128]; BW b; bw_init(&b, tmp, sizeof tmp);
    const vcpenc_params* p = &e->p;
    if (p->transform8x8) { bw_put(&b, 8, 100); bw_put(&b, 8, 0x00); }  /* High */
    else if (p->entropy) { bw_put(&b, 8, 77); bw_put(&b, 8, 0x40); }   /* Main (CABAC), constraint_set1 */
    else { bw_put(&b, 8, 66); bw_put(&b, 8, 0xC0); }                   /* Constrained Baseline: constraint_set0,1 */
    bw_put(&b, 8, level_idc_for(e->mbw, e->mbh, p->fps_num, p->fps_den));
    bw_ue(&b, 0);                   /* sps id */
    if (p->transform8x8) {          /* profile_idc 100: chroma format and bit depth fields */
        bw_ue(&b, 1);               /* chroma_format_idc 4:2:0 */
        bw_ue(&b, 0); bw_ue(&b, 0); /* bit_depth_luma/chroma_minus8 */
        bw_put(&b, 1, 0);           /* qpprime_y_zero_transform_bypass */
        bw_put(&b, 1, 0);           /* seq_scaling_matrix_present */
    }
    bw_ue(&b, 4);                   /* log2_max_frame_num_minus4 -> 8 bits */
    bw_ue(&b, 2);                   /* pic_order_cnt_type 2: output order == decode order */
    bw_ue(&b, 1);                   /* max_num_ref_frames */
    bw_put(&b, 1, 0);               /* gaps_in_frame_num_value_allowed */
    bw_ue(&b, e->mbw - 1);
    bw_ue(&b, e->mbh - 1);
    bw_put(&b, 1, 1);               /* frame_mbs_only */
    bw_put(&b, 1, 1);               /* direct_8x8_inference */
    int cr = e->cw - p->width, cb = e->ch - p->height;
    if (cr || cb) { bw_put(&b, 1, 1); bw_ue(&b, 0); bw_ue(&b, cr / 2); bw_ue(&b, 0); bw_ue(&b, cb / 2); }
    else bw_put(&b, 1, 0);
    bw_put(&b, 1, 1);               /* vui_parameters_present */
    bw_put(&b, 1, 0);               /* aspect_ratio_info_present */
    bw_put(&b, 1, 0);               /* overscan_info_present */
    bw_put(&b, 1, 0);               /* video_signal_type_present */
    bw_put(&b, 1, 0);               /* chroma_loc_info_present */
    bw_put(&b, 1, 1);               /* timing_info_present */
    bw_put32(&b, (uint32_t)p->fps_den);
    bw_put32(&b, (uint32_t)p->fps_num * 2);
    bw_put(&b, 1, 1);               /* fixed_frame_rate */
    bw_put(&b, 1, 0);               /* nal_hrd */
    bw_put(&b, 1, 0);               /* vcl_hrd */
    bw_put(&b, 1, 0);               /* pic_struct_present */
    bw_put(&b, 1, 1);               /* bitstream_restriction */
    bw_put(&b, 1, 1);               /* motion_vectors_over_pic_boundaries */
    bw_ue(&b, 0);                   /* max_bytes_per_pic_denom */
    bw_ue(&b, 0);                   /* max_bits_per_mb_denom */
    bw_ue(&b, 9);                   /* log2_max_mv_length_horizontal */
    bw_ue(&b, 9);                   /* log2_max_mv_length_vertical */
    bw_ue(&b, 0);                   /* max_num_reorder_frames */
    bw_ue(&b, 1);                   /* max_dec_frame_buffering */
    bw_trailing(&b);
    return nal_write(out, cap, 3, 7, tmp, b.pos);
}
static size_t write_pps(const Enc* e, uint8_t* out, size_t cap) {
    uint8_t tmp[64]; BW b; bw_init(&b, tmp, sizeof tmp);
    bw_ue(&b, 0); bw_ue(&b, 0);
    bw_put(&b, 1, e->p.entropy ? 1 : 0); /* entropy_coding_mode: 0 CAVLC, 1 CABAC */
    bw_put(&b, 1, 0);               /* bottom_field_pic_order_in_frame_present */
    bw_ue(&b, 0);                   /* num_slice_groups_minus1 */
    bw_ue(&b, 0); bw_ue(&b, 0);     /* num_ref_idx_l0/l1_default_active_minus1 */
    bw_put(&b, 1, 0);               /* weighted_pred */
    bw_put(&b, 2, 0);               /* weighted_bipred_idc */
    bw_se(&b, 0);                   /* pic_init_qp_minus26 */
    bw_se(&b, 0);                   /* pic_init_qs_minus26 */
    bw_se(&b, 0);                   /* chroma_qp_index_offset */
    bw_put(&b, 1, 1);               /* deblocking_filter_control_present */
    bw_put(&b, 1, 0);               /* constrained_intra_pred */
    bw_put(&b, 1, 0);               /* redundant_pic_cnt_present */
    if (e->p.transform8x8) {
        bw_put(&b, 1, 1);           /* transform_8x8_mode_flag */
        bw_put(&b, 1, 0);           /* pic_scaling_matrix_present */
        bw_se(&b, 0);               /* second_chroma_qp_index_offset */
    }
    bw_trailing(&b);
    return nal_write(out, cap, 3, 8, tmp, b.pos);
}
static void write_slice_header(const Enc* e, BW* b, int first_mb, int idr, int frame_num, int idr_id, int qp) {
    bw_ue(b, (unsigned)first_mb);
    bw_ue(b, idr ? 7 : 5);          /* slice_type: all slices of the picture are I / P */
    bw_ue(b, 0);                    /* pps id */
    bw_put(b, 8, (uint32_t)(frame_num & 255));
    if (idr) bw_ue(b, (unsigned)idr_id);
    if (!idr) {
        bw_put(b, 1, 0);            /* num_ref_idx_active_override */
        bw_put(b, 1, 0);            /* ref_pic_list_modification_flag_l0 */
    }
    if (idr) { bw_put(b, 1, 0); bw_put(b, 1, 0); } /* no_output_of_prior_pics, long_term_reference */
    else bw_put(b, 1, 0);           /* adaptive_ref_pic_marking_mode */
    if (e->p.entropy && !idr) bw_ue(b, 0); /* cabac_init_idc */
    bw_se(b, qp - 26);              /* slice_qp_delta */
    bw_ue(b, (unsigned)e->p.deblock_idc);
    if (e->p.deblock_idc != 1) { bw_se(b, 0); bw_se(b, 0); }
}

/* slice data for MB rows [r0,r1) */
static void write_slice_data(Enc* e, BW* b, int r0, int r1, int idr) {
    int skip_run = 0;
    for (int my = r0; my < r1; my++)
        for (int mx = 0; mx < e->mbw; mx++) {
            const MB* mb = &e->mbs[my * e->mbw + mx];
            if (!idr) {
                if (mb->type == VCP_MB_PSKIP) { skip_run++; continue; }
                bw_ue(b, (unsigned)skip_run); skip_run = 0;
            }
            if (mb->type == VCP_MB_I16) {
                int t = 1 + mb->i16_mode + 4 * (mb->cbp >> 4) + ((mb->cbp & 15) ? 12 : 0);
                bw_ue(b, (unsigned)(idr ? t : t + 5));
                bw_ue(b, mb->chroma_mode);
                bw_se(b, 0); /* mb_qp_delta */
            } else {
                bw_ue(b, 0); /* P_L0_16x16 */
                bw_se(b, mb->mvd[0]); bw_se(b, mb->mvd[1]);
                bw_ue(b, vcp_cbp_to_golomb_inter[mb->cbp]);
                if (e->p.transform8x8 && (mb->cbp & 15)) bw_put(b, 1, mb->t8x8);   /* transform_size_8x8_flag */
                if (mb->cbp) bw_se(b, 0);
            }
            cavlc_residual(b, e, mb, mx, my);
        }
    if (skip_run) bw_ue(b, (unsigned)skip_run);
    bw_trailing(b);
}

/* ------------------------------------------------------------------------------------ */
/* public entry points (ctypes)                                                            */
typedef struct orc_dump {
    int16_t* mv_prepass; /* [nframes][nmb][2] full-pel * 4 */
    int16_t* mv_final;   /* [nframes][nmb][2] quarter-pel */
    uint8_t* mb_type;    /* [nframes][nmb] */
    uint8_t* cbp;        /* [nframes][nmb] */
} orc_dump;

static void store_recon(const Enc* e, const Frame* f, uint8_t* dst) {
    int w = e->p.width, h = e->p.height, cw = (w + 1) / 2, chh = (h + 1) / 2;
    for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * w, f->y + (size_t)y * f->ys, w);
    dst += (size_t)w * h;
    for (int y = 0; y < chh; y++) memcpy(dst + (size_t)y * cw, f->u + (size_t)y * f->cs, cw);
    dst += (size_t)cw * chh;
    for (int y = 0; y < chh; y++) memcpy(dst + (size_t)y * cw, f->v + (size_t)y * f->cs, cw);
}

int orc_encode(const vcpenc_params* p, const uint8_t* frames, int nframes, uint8_t* out, size_t out_cap,
               size_t* out_len, vcpenc_frame_info* info, uint8_t* recon, orc_dump* dump) {
    Enc E; Enc* e = &E;
    memset(e, 0, sizeof *e);
    e->p = *p;
    if (e->p.slices == 0) e->p.slices = vcp_auto_slices((p->height + 15) / 16, p->entropy);
    p = &e->p;
    if (p->width < 16 || p->height < 16 || (p->width & 1) || (p->height & 1) || p->gop < 1 || p->slices < 0) return VCPENC_E_ARGS;
    if (p->entropy < 0 || p->entropy > 1 || p->codec != VCPENC_CODEC_H264 || p->in_fmt < 0 || p->in_fmt > VCPENC_FMT_BGR24) return VCPENC_E_ARGS;
    if (p->rc_mode == VCPENC_RC_ABR && p->bitrate <= 0) return VCPENC_E_ARGS;
    e->mbw = (p->width + 15) / 16; e->mbh = (p->height + 15) / 16; e->nmb = e->mbw * e->mbh;
    if (p->slices > e->mbh) return VCPENC_E_ARGS;
    e->cw = 16 * e->mbw; e->ch = 16 * e->mbh;
    int rc = VCPENC_E_INTERNAL;
    if (frame_alloc(&e->cur, e->cw, e->ch) || frame_alloc(&e->prev_orig, e->cw, e->ch) ||
        frame_alloc(&e->recon[0], e->cw, e->ch) || frame_alloc(&e->recon[1], e->cw, e->ch) ||
        half_alloc(&e->hcur, e->cw, e->ch) || half_alloc(&e->hprev, e->cw, e->ch)) goto done;
    e->mbs = (MB*)calloc((size_t)e->nmb, sizeof(MB));
    e->mvfp = (int16_t*)calloc((size_t)e->nmb * 2, sizeof(int16_t));
    e->rbsp_cap = (size_t)e->nmb * 512 + 4096;
    e->rbsp = (uint8_t*)malloc(e->rbsp_cap);
    if (!e->mbs || !e->mvfp || !e->rbsp) goto done;

    size_t fsz = (size_t)p->width * p->height + 2 * (size_t)((p->width + 1) / 2) * ((p->height + 1) / 2);
    const int iw = p->in_width > 0 ? p->in_width : p->width, ih = p->in_height > 0 ? p->in_height : p->height;
    const size_t in_fsz = (size_t)vcp_in_frame_bytes(p->in_fmt, iw, ih);
    const int need_k1 = p->in_fmt != VCPENC_FMT_YUV420P || iw != p->width || ih != p->height;
    uint8_t* k1_a = need_k1 ? (uint8_t*)malloc((size_t)iw * ih * 3 / 2 + iw + ih + 16) : NULL;
    uint8_t* k1_b = need_k1 ? (uint8_t*)malloc(fsz + 16) : NULL;
    if (need_k1 && (!k1_a || !k1_b)) { free(k1_a); free(k1_b); goto done; }
    size_t o = 0;
    int ri = 0, idr_count = p->first_gop;
    /* rate control state of the current GOP (vcp_algo.h) */
    const int abr = p->rc_mode == VCPENC_RC_ABR;
    const int vbv = vcp_rc_has_vbv(p->maxrate, p->bufsize, p->fps_num, p->fps_den);
    const int rc_fb = abr || vbv;   /* per-picture QP feedback on */
    const int rc_qp0 = abr ? vcp_rc_initial_qp(vcp_rc_eff_bitrate(p->bitrate, p->maxrate), p->fps_num, p->fps_den, p->width, p->height) : 0;
    unsigned long long rc_cum = 0;
    long long rc_full = 0;
    int rc_qp_next[2] = {0, 0}; /* QP decided for pictures t+1, t+2 */
    for (int n = 0; n < nframes; n++) {
        int t = n % p->gop, idr = t == 0;
        int qp = idr ? p->qp_i : p->qp_p;
        if (rc_fb) {
            if (idr) {
                rc_qp_next[0] = rc_qp_next[1] = abr ? rc_qp0 : p->qp_p;
                if (abr) { qp = rc_qp0 - VCP_RC_QP_I_OFFSET; if (qp < 0) qp = 0; }
            } else { qp = rc_qp_next[0]; rc_qp_next[0] = rc_qp_next[1]; }
        }
        /* K1 */
        { Frame tf = e->prev_orig; e->prev_orig = e->cur; e->cur = tf; Half th = e->hprev; e->hprev = e->hcur; e->hcur = th; }
        {
            const uint8_t* fin = frames + (size_t)n * in_fsz;
            if (need_k1) {
                to_yuv420p(p->in_fmt, fin, iw, ih, k1_a);
                if (iw != p->width || ih != p->height) { scale_yuv420p(k1_a, iw, ih, k1_b, p->width, p->height); fin = k1_b; }
                else fin = k1_a;
            }
            frame_load_yuv420p(&e->cur, fin, p->width, p->height);
        }
        half_build(&e->hcur, &e->cur);
        Frame* rec = &e->recon[ri]; const Frame* ref = &e->recon[ri ^ 1];
        if (idr) {
            for (int my = 0; my < e->mbh; my++)
                for (int mx = 0; mx < e->mbw; mx++) encode_i_mb(e, rec, mx, my, qp);
            if (dump) {
                if (dump->mv_prepass) memset(dump->mv_prepass + (size_t)n * e->nmb * 2, 0, (size_t)e->nmb * 4);
            }
        } else {
            me_prepass(e);
            for (int i = 0; i < e->nmb; i++) {
                me_refine_mb(e, ref, i % e->mbw, i / e->mbw, qp, e->mbs[i].mv);
                encode_p_mb(e, ref, rec, i % e->mbw, i / e->mbw, qp);
            }
            /* intra macroblocks of the P picture: they predict from the reconstruction of their
             * neighbours (inter ones are complete, intra ones precede in raster order) */
            for (int i = 0; i < e->nmb; i++)
                if (e->mbs[i].type == VCP_MB_I16) encode_i_mb(e, rec, i % e->mbw, i / e->mbw, qp);
            /* post-hoc: predictors, differences, skip */
            for (int i = 0; i < e->nmb; i++) {
                MB* mb = &e->mbs[i];
                int sx, sy, px, py;
                mv_pskip(e, i % e->mbw, i / e->mbw, &sx, &sy);
                mvp16(e, i % e->mbw, i / e->mbw, &px, &py);
                if (mb->type == VCP_MB_P16 && !mb->cbp && mb->mv[0] == sx && mb->mv[1] == sy) mb->type = VCP_MB_PSKIP;
                mb->mvd[0] = (int16_t)(mb->mv[0] - px); mb->mvd[1] = (int16_t)(mb->mv[1] - py);
            }
            if (dump && dump->mv_prepass)
                for (int i = 0; i < 2 * e->nmb; i++) dump->mv_prepass[(size_t)n * e->nmb * 2 + i] = (int16_t)(4 * e->mvfp[i]);
        }
        if (dump) for (int i = 0; i < e->nmb; i++) {
            if (dump->mv_final) { dump->mv_final[((size_t)n * e->nmb + i) * 2] = e->mbs[i].mv[0]; dump->mv_final[((size_t)n * e->nmb + i) * 2 + 1] = e->mbs[i].mv[1]; }
            if (dump->mb_type) dump->mb_type[(size_t)n * e->nmb + i] = e->mbs[i].type;
            if (dump->cbp) dump->cbp[(size_t)n * e->nmb + i] = e->mbs[i].cbp;
        }
        /* K5 + NAL */
        size_t au0 = o;
        if (idr) {
            size_t k = write_sps(e, out + o, out_cap - o); if (!k) { rc = VCPENC_E_OVERFLOW; goto done; } o += k;
            k = write_pps(e, out + o, out_cap - o); if (!k) { rc = VCPENC_E_OVERFLOW; goto done; } o += k;
        }
        unsigned long long frame_bits = 0;
        for (int s = 0; s < p->slices; s++) {
            int r0 = slice_first_row(e, s), r1 = s + 1 < p->slices ? slice_first_row(e, s + 1) : e->mbh;
            BW b; bw_init(&b, e->rbsp, e->rbsp_cap);
            write_slice_header(e, &b, r0 * e->mbw, idr, t, idr_count & 1, qp);
            unsigned long long nb = 0;
            if (p->entropy) write_slice_data_cabac(e, &b, r0, r1, idr, qp, &nb); else write_slice_data(e, &b, r0, r1, idr);
            if (b.overflow) { rc = VCPENC_E_OVERFLOW; goto done; }
            size_t k = nal_write(out + o, out_cap - o, idr ? 3 : 2, idr ? 5 : 1, e->rbsp, b.pos);
            if (!k) { rc = VCPENC_E_OVERFLOW; goto done; }
            o += k;
            /* CABAC runs after the whole recon chain on the device: rate control sees the bin
             * count scaled by VCP_CABAC_BITS_PER_BIN_Q4/16 instead of the final bits */
            frame_bits += p->entropy ? (nb * VCP_CABAC_BITS_PER_BIN_Q4) >> 4 : (unsigned long long)b.pos * 8;
        }
        if (rc_fb) {
            /* feedback lands two pictures later (entropy coding runs beside the recon chain) */
            int gop_len = p->gop;
            int g0 = n - t;
            if (g0 + gop_len > nframes) gop_len = nframes - g0;
            rc_qp_next[1] = vcp_rc_picture(abr, rc_qp0, p->qp_p, abr ? vcp_rc_gop_budget(p->bitrate, p->maxrate, p->fps_num, p->fps_den, gop_len) : 0,
                                           vbv ? vcp_vbv_rate(p->maxrate, p->fps_num, p->fps_den) : 0, vbv ? p->bufsize : 0,
                                           qp, rc_qp_next[0], idr, frame_bits, t, gop_len, &rc_cum, &rc_full);
        }
        if (idr) idr_count++;
        if (info) { info[n].offset = au0; info[n].size = (uint32_t)(o - au0); info[n].is_idr = (uint8_t)idr; info[n].qp = (uint8_t)qp; }
        /* K4 */
        deblock_frame(e, rec, qp);
        frame_pad(rec);
        if (recon) store_recon(e, rec, recon + (size_t)n * fsz);
        ri ^= 1;
    }
    *out_len = o;
    rc = VCPENC_OK;
    free(k1_a); free(k1_b);
done:
    frame_free(&e->cur); frame_free(&e->prev_orig); frame_free(&e->recon[0]); frame_free(&e->recon[1]);
    free(e->hcur.buf); free(e->hprev.buf); free(e->mbs); free(e->mvfp); free(e->rbsp);
    return rc;
}

/* small known-answer taps for unit tests */
void orc_fdct4(const int* d, int* w) { fdct4(d, w); }
void orc_idct4(const int* c, int* r) { idct4(c, r); }
int orc_quant4x4(const int* w, int qp, int intra, int first, int16_t* lv) { return quant4x4(w, qp, intra, first, lv); }
void orc_dequant4x4(const int16_t* lv, int qp, int first, int* c) { memset(c, 0, 64); dequant4x4(lv, qp, first, c); }
int orc_cavlc_block(const int16_t* c, int n, int nC, uint8_t* out, int cap) {
    BW b; bw_init(&b, out, (size_t)cap);
    cavlc_block(&b, c, n, nC);
    int bits = (int)bw_bits(&b);
    if (b.nbits) bw_put(&b, 8 - b.nbits, 0);
    return bits;
}
int orc_ue_bits(unsigned k, uint8_t* out) { BW b; bw_init(&b, out, 8); bw_ue(&b, k); int n = (int)bw_bits(&b); if (b.nbits) bw_put(&b, 8 - b.nbits, 0); return n; }
int orc_se_bits(int v, uint8_t* out) { BW b; bw_init(&b, out, 8); bw_se(&b, v); int n = (int)bw_bits(&b); if (b.nbits) bw_put(&b, 8 - b.nbits, 0); return n; }
int orc_luma_qpel(const uint8_t* plane, int stride, int ix, int iy, int fx, int fy) { return luma_qpel(plane, stride, ix, iy, fx, fy); }
int orc_lambda(int qp) { return vcp_lambda(qp); }
int orc_decimate_score(unsigned long long mask, int big, int is8x8) { return vcp_decimate_score(mask, big, is8x8); }
void orc_fdct8(const int* d, int* w) { fdct8(d, w); }
void orc_roundtrip8(const int* d, int qp, int* r) {
    int w[64], c[64]; int16_t lv[64];
    fdct8(d, w); quant8x8(w, qp, 0, lv); dequant8x8(lv, qp, c); idct8(c, r);
}

/* HEVC oracle (config #4): shares the helpers above */
#include "hevc_oracle.inc.c"
