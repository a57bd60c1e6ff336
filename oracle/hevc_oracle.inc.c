/* CPU oracle of the HEVC path (BASELINE config #4) -- TEST INFRASTRUCTURE ONLY.
 * Included at the end of h264_oracle.c so that it shares the frame / motion-search / bit-writer /
 * arithmetic-coder helpers.  This
 * file is pinned the same way as the H.264 oracle, by the FFmpeg `hevc` decoder reproducing the
 * reconstruction bit-exactly; the CUDA path (csrc/k6_hevc.cu, k5_cabac.cu) must match it byte for byte.
 *
 * No reference file to follow (the reference shells out to libx265 / hevc_nvenc through ffmpeg,
 * /root/reference/internal/config/config.go:47-50); the arithmetic restates ITU-T H.265.
 *
 * Stream structure (every choice is GPU-friendly and mirrors the H.264 path's parallelism):
 *   - Main profile, 8-bit 4:2:0, closed GOPs (IDR_W_RADL + TRAIL_R), one reference, POC = decode order
 *   - CTB = CU = 16x16 (no split flags), coded in raster order; slices = whole CTB rows
 *   - transform blocks: luma 8x8 (the 16x16 root is split because MaxTb = 8), chroma 4x4; DCT only
 *   - intra: DC prediction per transform block (mode signalled through the MPM list), chroma derived; every CU of an
 *     IDR picture, and CUs of P pictures where the refine finds intra cheaper (scene cuts)
 *   - inter: one 16x16 PU, full-sample luma vectors (chroma lands on half samples: 4-tap filter); with
 *     params.hevc_subpel >= 1 also the 8 half-sample neighbours (8-tap luma filter), >= 2 the 8 quarter-sample
 *     neighbours of that (7/8-tap filters; 2: ranked by a half-sample-average proxy, 3: by the exact prediction),
 *     AMVP with spatial candidates, merge (1 candidate) / skip
 *   - CABAC (same engine as H.264 9.3.4.2; HEVC context tables 9-5..9-37), no sign hiding
 *   - in-loop deblocking (all vertical edges, then all horizontal ones; not across slices) unless deblock_idc = 1;
 *     SAO disabled in the SPS; constant QP or the bitrate model of the H.264 path
 */

#include "../video_codec_pipeline_b200/csrc/hevc_tables.h"

/* ---- per coding unit ----------------------------------------------------------------------- */
enum { HCU_INTRA = 0, HCU_INTER = 1, HCU_SKIP = 2 };
typedef struct {
    uint8_t type;            /* HCU_* */
    uint8_t merge;           /* inter: merge_flag */
    uint8_t mvp_idx;
    uint8_t imode;           /* intra: luma prediction mode (0 planar, 1 DC, 10 horizontal, 26 vertical); chroma derives it */
    uint8_t cbf_y[4], cbf_cb[4], cbf_cr[4];   /* per 8x8 transform unit, z-order */
    int16_t mv[2];           /* quarter-sample units, multiples of 4 */
    int16_t mvd[2];
    int16_t lv_y[4][64], lv_cb[4][16], lv_cr[4][16];   /* raster inside the block */
} HCU;

/* ---- CABAC ------------------------------------------------------------------------------ */
static void hevc_cabac_init(Cabac* c, BW* b, int init_type, int qp) {
    c->b = b; c->low = 0; c->range = 510; c->outstanding = 0; c->first = 1; c->nbins = 0;
    for (int i = 0; i < HC_NCTX; i++) {
        int v = hevc_init_values[init_type][i];
        int slope = v >> 4, offs = v & 15;
        int m = slope * 5 - 45, n = (offs << 3) - 16;
        int pre = vcp_clip3(1, 126, ((m * vcp_clip3(0, 51, qp)) >> 4) + n);
        c->state[i] = pre <= 63 ? (uint8_t)((63 - pre) << 1) : (uint8_t)(((pre - 64) << 1) | 1);
    }
}

/* ---- transforms (8.6.4.2) and quantisation (8.6.3) -------------------------------------------- */
static void hevc_fwd(const int* d, int* w, int n) {   /* encoder side: W = M * D * M^T with HM's stage shifts */
    int t[64];
    const int s1 = (n == 8 ? 3 : 2) - 1, s2 = (n == 8 ? 3 : 2) + 6;
    for (int k = 0; k < n; k++)            /* rows of D through M: t[k][y] = sum_x M[k][x] * d[y][x] */
        for (int y = 0; y < n; y++) {
            int acc = 0;
            for (int x = 0; x < n; x++) acc += (n == 8 ? hevc_dct8[k][x] : hevc_dct4[k][x]) * d[y * n + x];
            t[k * n + y] = (acc + (s1 ? 1 << (s1 - 1) : 0)) >> s1;
        }
    for (int l = 0; l < n; l++)            /* w[l][k] = sum_y M[l][y] * t[k][y] */
        for (int k = 0; k < n; k++) {
            int acc = 0;
            for (int y = 0; y < n; y++) acc += (n == 8 ? hevc_dct8[l][y] : hevc_dct4[l][y]) * t[k * n + y];
            w[l * n + k] = (acc + (1 << (s2 - 1))) >> s2;
        }
}
static void hevc_inv(const int* c, int* r, int n) {   /* normative: columns first (>>7, clipped to 16 bit), then rows (>>12) */
    int g[64];
    for (int x = 0; x < n; x++)
        for (int y = 0; y < n; y++) {
            int acc = 0;
            for (int k = 0; k < n; k++) acc += (n == 8 ? hevc_dct8[k][y] : hevc_dct4[k][y]) * c[k * n + x];
            g[y * n + x] = vcp_clip3(-32768, 32767, (acc + 64) >> 7);
        }
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {
            int acc = 0;
            for (int k = 0; k < n; k++) acc += (n == 8 ? hevc_dct8[k][x] : hevc_dct4[k][x]) * g[y * n + k];
            r[y * n + x] = (acc + 2048) >> 12;
        }
}
static int hevc_quant(const int* w, int n, int qp, int intra, int16_t* lv) {
    const int log2n = n == 8 ? 3 : 2, tshift = 15 - 8 - log2n, qbits = 14 + qp / 6 + tshift;
    const long long offs = (long long)(intra ? 171 : 85) << (qbits - 9);
    int nz = 0;
    for (int i = 0; i < n * n; i++) {
        long long a = w[i] < 0 ? -(long long)w[i] : w[i];
        long long l = (a * hevc_quant_scale[qp % 6] + offs) >> qbits;
        if (l > 32767) l = 32767;
        lv[i] = (int16_t)(w[i] < 0 ? -l : l);
        nz += l != 0;
    }
    return nz;
}
static void hevc_dequant(const int16_t* lv, int n, int qp, int* c) {
    const int log2n = n == 8 ? 3 : 2, bdshift = 8 + log2n + 10 - 15;
    for (int i = 0; i < n * n; i++) {
        long long v = ((long long)lv[i] * 16 * hevc_level_scale[qp % 6]) << (qp / 6);
        c[i] = vcp_clip3(-32768, 32767, (int)((v + (1 << (bdshift - 1))) >> bdshift));
    }
}
static int hevc_chroma_qp(int qp) { return qp < 30 ? qp : qp > 43 ? qp - 6 : hevc_qpc_tab[qp - 30]; }

/* scan position -> (x, y) inside a 4x4 block / sub-block index -> (xs, ys) inside an 8x8 block, per scanIdx
 * (6.5.3 up-right diagonal, 6.5.4 horizontal, 6.5.5 vertical) */
static void hevc_scan4(int scan, int p, int* x, int* y) {
    if (scan == 0) { *x = hevc_diag4_x[p]; *y = hevc_diag4_y[p]; }
    else if (scan == 1) { *x = p & 3; *y = p >> 2; }
    else { *x = p >> 2; *y = p & 3; }
}
static void hevc_scan2(int scan, int i, int* x, int* y) {
    if (scan == 0) { *x = hevc_diag2_x[i]; *y = hevc_diag2_y[i]; }
    else if (scan == 1) { *x = i & 1; *y = i >> 1; }
    else { *x = i >> 1; *y = i & 1; }
}
/* scanIdx of an intra block (7.4.9.11): horizontal-ish modes scan vertically, vertical-ish modes horizontally */
static int hevc_scan_idx(int intra, int mode) { return !intra ? 0 : (mode >= 6 && mode <= 14) ? 2 : (mode >= 22 && mode <= 30) ? 1 : 0; }

/* ---- residual_coding (7.3.8.11, 9.3.4.2.x) for 8x8 luma / 4x4 chroma ------------------------------ */
static void hevc_residual(Cabac* cb, const int16_t* lv /* raster */, int log2n, int cidx, int scan) {
    const int n = 1 << log2n, nsb = n == 8 ? 4 : 1;
    /* coefficients in coding (scan) order: sub-block i, position p */
    int pos_x[64], pos_y[64], coef[64], ncoef = n * n, last = -1;
    for (int i = 0; i < nsb; i++)
        for (int p = 0; p < 16; p++) {
            int sx = 0, sy = 0, x, y;
            if (n == 8) hevc_scan2(scan, i, &sx, &sy);
            hevc_scan4(scan, p, &x, &y);
            x += 4 * sx; y += 4 * sy;
            pos_x[16 * i + p] = x; pos_y[16 * i + p] = y; coef[16 * i + p] = lv[y * n + x];
            if (coef[16 * i + p]) last = 16 * i + p;
        }
    (void)ncoef;
    /* last significant position: prefix (context coded, truncated unary) + suffix (bypass) */
    {
        int lx = pos_x[last], ly = pos_y[last];
        if (scan == 2) { const int t = lx; lx = ly; ly = t; }   /* vertical scan: the position is coded transposed */
        int off, shift;
        if (cidx == 0) { off = 3 * (log2n - 2) + ((log2n - 1) >> 2); shift = (log2n + 1) >> 2; }
        else { off = 15; shift = log2n - 2; }
        const int cmax = (log2n << 1) - 1;
        for (int comp = 0; comp < 2; comp++) {
            int v = comp ? ly : lx;
            int prefix = v < 4 ? v : (v < 6 ? 4 : 5);    /* positions 0..7 */
            for (int i = 0; i < prefix; i++) cabac_encode(cb, (comp ? HC_LAST_Y : HC_LAST_X) + off + (i >> shift), 1);
            if (prefix < cmax) cabac_encode(cb, (comp ? HC_LAST_Y : HC_LAST_X) + off + (prefix >> shift), 0);
        }
        for (int comp = 0; comp < 2; comp++) {
            int v = comp ? ly : lx;
            if (v >= 4) cabac_bypass(cb, v & 1);          /* one suffix bit for prefixes 4 and 5 */
        }
    }
    int csbf[2][2] = {{0, 0}, {0, 0}};
    int prev_gt1_zero = 0;    /* the previous sub-block ended with greater1Ctx == 0 */
    int first_sb = 1;
    for (int i = last >> 4; i >= 0; i--) {
        int xs = 0, ys = 0;
        if (n == 8) hevc_scan2(scan, i, &xs, &ys);
        const int* cf = coef + 16 * i;
        int any = 0;
        for (int p = 0; p < 16; p++) any |= cf[p] != 0;
        const int right = xs + 1 < 2 && n == 8 ? csbf[ys][xs + 1] : 0, below = ys + 1 < 2 && n == 8 ? csbf[ys + 1][xs] : 0;
        int infer_dc = 0;
        if (i < (last >> 4) && i > 0) {
            cabac_encode(cb, HC_CSBF + (cidx ? 2 : 0) + ((right | below) ? 1 : 0), any);
            infer_dc = 1;
        } else any = 1;   /* first and last sub-block: inferred 1 */
        csbf[ys][xs] = any;
        if (!any) continue;
        /* significance map */
        const int start = (i == (last >> 4)) ? (last & 15) - 1 : 15;
        int sig[16], nsig = 0;
        for (int p = 0; p < 16; p++) sig[p] = 0;
        if (i == (last >> 4)) { sig[last & 15] = 1; infer_dc = 0; }
        for (int p = start; p >= 0; p--) {
            const int s = cf[p] != 0;
            if (p > 0 || !infer_dc || 1) {
                /* the DC of a coded sub-block is inferred when every other coefficient of it was zero */
                int coded = 1;
                if (p == 0 && infer_dc) {
                    int others = 0;
                    for (int q = 1; q < 16; q++) others |= sig[q];
                    if (!others) coded = 0;
                }
                if (coded) {
                    int xp, yp, sc;
                    hevc_scan4(scan, p, &xp, &yp);
                    if (log2n == 2) sc = hevc_sig_ctx_map4[(yp << 2) + xp];
                    else if (xs == 0 && ys == 0 && p == 0) sc = 0;
                    else {
                        const int pat = right | (below << 1);
                        if (pat == 0) sc = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
                        else if (pat == 1) sc = yp == 0 ? 2 : yp == 1 ? 1 : 0;
                        else if (pat == 2) sc = xp == 0 ? 2 : xp == 1 ? 1 : 0;
                        else sc = 2;
                        if (cidx == 0) { if (xs || ys) sc += 3; sc += scan == 0 ? 9 : 15; }   /* log2 == 3 */
                        else sc += 9;
                    }
                    cabac_encode(cb, HC_SIG + (cidx == 0 ? sc : 27 + sc), s);
                }
            }
            sig[p] = s;
        }
        /* levels of this sub-block, from the highest scan position down */
        int idx[16];
        for (int p = 15; p >= 0; p--) if (sig[p]) idx[nsig++] = p;
        int ctxset = (i == 0 || cidx > 0) ? 0 : 2;
        if (!first_sb && prev_gt1_zero) ctxset++;
        first_sb = 0;
        int g1ctx = 1, g1flag[8], first_g2 = -1, ng1 = nsig < 8 ? nsig : 8;
        for (int k = 0; k < ng1; k++) {
            const int a = abs(cf[idx[k]]);
            g1flag[k] = a > 1;
            cabac_encode(cb, HC_GT1 + (cidx ? 16 : 0) + ctxset * 4 + g1ctx, g1flag[k]);
            if (g1flag[k]) { g1ctx = 0; if (first_g2 < 0) first_g2 = k; }
            else if (g1ctx > 0 && g1ctx < 3) g1ctx++;
        }
        prev_gt1_zero = g1ctx == 0;
        int g2flag = 0;
        if (first_g2 >= 0) {
            g2flag = abs(cf[idx[first_g2]]) > 2;
            cabac_encode(cb, HC_GT2 + (cidx ? 4 : 0) + ctxset, g2flag);
        }
        for (int k = 0; k < nsig; k++) cabac_bypass(cb, cf[idx[k]] < 0);
        int rice = 0;
        for (int k = 0; k < nsig; k++) {
            const int a = abs(cf[idx[k]]);
            const int base = k < 8 ? 1 + g1flag[k] + (k == first_g2 ? g2flag : 0) : 1;
            const int thresh = k < 8 ? (k == first_g2 ? 3 : 2) : 1;
            if (base == thresh) {
                int rem = a - base;
                /* coeff_abs_level_remaining (9.3.3.11): prefix up to 3 in unary with rice suffix, then escape */
                if (rem < (3 << rice)) {
                    int len = rem >> rice;
                    for (int q = 0; q < len; q++) cabac_bypass(cb, 1);
                    cabac_bypass(cb, 0);
                    for (int q = rice - 1; q >= 0; q--) cabac_bypass(cb, (rem >> q) & 1);
                } else {
                    int len = rice, v = rem - (3 << rice);
                    while (v >= (1 << len)) { v -= 1 << len; len++; }
                    for (int q = 0; q < 3 + len + 1 - rice - 1; q++) cabac_bypass(cb, 1);
                    cabac_bypass(cb, 0);
                    for (int q = len - 1; q >= 0; q--) cabac_bypass(cb, (v >> q) & 1);
                }
                if (a > 3 * (1 << rice) && rice < 4) rice++;
            }
        }
    }
}

/* ---- encoder state ----------------------------------------------------------------------------- */
typedef struct {
    Enc* e;
    HCU* cus;
    int qp, qpc, idr;
    Frame* rec; const Frame* ref;
    uint8_t* sao_type;        /* per CTB: 0 off, 1 + edge class */
    int8_t* sao_off;          /* per CTB: four offsets (categories 1..4) */
    uint8_t* sao_tmp;         /* copy of the deblocked luma picture while offsets are applied */
} HEnc;

/* availability of the 4x4 luma unit at (x,y) for a block at (xc,yc) decoded in z-order inside 16x16 CTBs (6.4.1) */
static int hevc_avail(const Enc* e, int xc, int yc, int x, int y) {
    if (x < 0 || y < 0 || x >= e->cw || y >= e->ch) return 0;
    const int cx = xc >> 4, cy = yc >> 4, nx = x >> 4, ny = y >> 4;
    const int row0 = slice_first_row(e, slice_of_row(e, cy));
    if (ny < row0) return 0;                          /* other slice */
    if (ny < cy) return 1;
    if (ny > cy) return 0;
    if (nx < cx) return 1;
    if (nx > cx) return 0;
    const int zc = ((yc >> 3) & 1) * 2 + ((xc >> 3) & 1), zn = ((y >> 3) & 1) * 2 + ((x >> 3) & 1);
    return zn < zc;
}

/* Intra prediction of one transform block (8.4.4.2): reference substitution (8.4.4.2.2), [1 2 1] smoothing of the
 * references where the mode asks for it (8.4.4.2.3: luma 8x8 planar), then planar (0), DC (1), horizontal (10) or
 * vertical (26) with the luma edge filters of DC / horizontal / vertical */
static void hevc_intra_pred(const HEnc* h, int cidx, int xl, int yl /* luma position of the TU */, int n, int mode, uint8_t* pred) {
    const Enc* e = h->e;
    const Frame* f = h->rec;
    const uint8_t* plane = cidx == 0 ? f->y : cidx == 1 ? f->u : f->v;
    const int stride = cidx == 0 ? f->ys : f->cs, sh = cidx ? 1 : 0;
    const int x0 = xl >> sh, y0 = yl >> sh;
    /* reference samples: ref[0] = bottom of the extended left column ... ref[2n] = corner ... ref[4n] = end of the top row */
    int ref[33], av[33];
    for (int k = 0; k <= 4 * n; k++) {
        int x, y;
        if (k < 2 * n) { x = -1; y = 2 * n - 1 - k; } else if (k == 2 * n) { x = -1; y = -1; } else { x = k - 2 * n - 1; y = -1; }
        av[k] = hevc_avail(e, xl, yl, xl + (x << sh), yl + (y << sh));
        ref[k] = av[k] ? plane[(size_t)(y0 + y) * stride + x0 + x] : 0;
    }
    int any = 0;
    for (int k = 0; k <= 4 * n; k++) any |= av[k];
    if (!any) for (int k = 0; k <= 4 * n; k++) ref[k] = 128;
    else {
        if (!av[0]) { int k = 1; while (!av[k]) k++; ref[0] = ref[k]; }
        for (int k = 1; k <= 4 * n; k++) if (!av[k]) ref[k] = ref[k - 1];
    }
    if (cidx == 0 && n == 8 && mode == 0) {           /* filterFlag: min(|mode - 26|, |mode - 10|) > 7 -- of our modes only planar */
        int fl[33];
        fl[0] = ref[0]; fl[4 * n] = ref[4 * n];
        for (int k = 1; k < 4 * n; k++) fl[k] = (ref[k - 1] + 2 * ref[k] + ref[k + 1] + 2) >> 2;
        for (int k = 0; k <= 4 * n; k++) ref[k] = fl[k];
    }
    const int* left = ref + 2 * n - 1;   /* left[-y] = p[-1][y] */
    const int* top = ref + 2 * n + 1;    /* top[x] = p[x][-1] */
    const int corner = ref[2 * n], lg = n == 8 ? 3 : 2;
    if (mode == 0) {                                   /* planar (8.4.4.2.4) */
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++)
                pred[y * n + x] = (uint8_t)(((n - 1 - x) * left[-y] + (x + 1) * top[n] + (n - 1 - y) * top[x] + (y + 1) * left[-n] + n) >> (lg + 1));
        return;
    }
    if (mode == 26) {                                  /* vertical (8.4.4.2.6, intraPredAngle 0) */
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) pred[y * n + x] = (uint8_t)top[x];
        if (cidx == 0) for (int y = 0; y < n; y++) pred[y * n] = (uint8_t)vcp_clip255(top[0] + ((left[-y] - corner) >> 1));
        return;
    }
    if (mode == 10) {                                  /* horizontal */
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) pred[y * n + x] = (uint8_t)left[-y];
        if (cidx == 0) for (int x = 0; x < n; x++) pred[x] = (uint8_t)vcp_clip255(left[0] + ((top[x] - corner) >> 1));
        return;
    }
    int sum = n;
    for (int i = 0; i < n; i++) sum += top[i] + left[-i];
    const int dc = sum >> ((n == 8 ? 3 : 2) + 1);
    for (int i = 0; i < n * n; i++) pred[i] = (uint8_t)dc;
    if (cidx == 0) {
        pred[0] = (uint8_t)((left[0] + 2 * dc + top[0] + 2) >> 2);
        for (int x = 1; x < n; x++) pred[x] = (uint8_t)((top[x] + 3 * dc + 2) >> 2);
        for (int y = 1; y < n; y++) pred[y * n] = (uint8_t)((left[-y] + 3 * dc + 2) >> 2);
    }
}

/* residual of one transform block: src - pred -> levels, recon written in place */
static int hevc_code_block(const HEnc* h, int cidx, int x0, int y0, int n, const uint8_t* pred, int intra, int16_t* lv) {
    const Enc* e = h->e;
    const uint8_t* sp = cidx == 0 ? e->cur.y : cidx == 1 ? e->cur.u : e->cur.v;
    uint8_t* rp = cidx == 0 ? h->rec->y : cidx == 1 ? h->rec->u : h->rec->v;
    const int ss = cidx == 0 ? e->cur.ys : e->cur.cs, rs = cidx == 0 ? h->rec->ys : h->rec->cs;
    int d[64], w[64], c[64], r[64];
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) d[y * n + x] = sp[(size_t)(y0 + y) * ss + x0 + x] - pred[y * n + x];
    hevc_fwd(d, w, n);
    const int nz = hevc_quant(w, n, cidx ? h->qpc : h->qp, intra, lv);
    if (nz) { hevc_dequant(lv, n, cidx ? h->qpc : h->qp, c); hevc_inv(c, r, n); }
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
            rp[(size_t)(y0 + y) * rs + x0 + x] = (uint8_t)vcp_clip255(pred[y * n + x] + (nz ? r[y * n + x] : 0));
    return nz != 0;
}

static const HCU* hevc_nb_raw(const HEnc* h, int cx, int cy) { return cx > 0 ? &h->cus[cy * h->e->mbw + cx - 1] : NULL; }
/* most probable modes of a CU (8.4.2): the CU above always sits in another CTB row (CTB = CU), so candidate B is DC */
static void hevc_mpm_list(const HEnc* h, int cx, int cy, int list[3]) {
    const HCU* l = hevc_nb_raw(h, cx, cy);
    const int a = l && l->type == HCU_INTRA ? l->imode : 1;
    if (a == 1) { list[0] = 0; list[1] = 1; list[2] = 26; return; }
    list[0] = a; list[1] = 1; list[2] = a != 0 ? 0 : 26;
}
static int hevc_mode_bins(const int list[3], int mode) {      /* prev_intra_luma_pred_flag + mpm_idx / rem_intra_luma_pred_mode */
    return mode == list[0] ? 2 : (mode == list[1] || mode == list[2]) ? 3 : 6;
}
/* Mode of an intra CU when params.hevc_intra_modes: SAD of a 16x16 prediction made from the CU's reconstructed top row,
 * left column and corner (missing side: the other side's first sample, both missing: 128; planar takes top[15] /
 * left[15] for the two far corners) + lambda x mode bins.  A decision rule, not a normative process. */
static int hevc_choose_intra_mode(const HEnc* h, int cx, int cy) {
    const Enc* e = h->e;
    const Frame* f = h->rec;
    const int row0 = slice_first_row(e, slice_of_row(e, cy));
    const int aL = cx > 0, aT = cy > row0;
    const uint8_t* r = f->y + (size_t)(16 * cy) * f->ys + 16 * cx;
    const uint8_t* c = e->cur.y + (size_t)(16 * cy) * e->cur.ys + 16 * cx;
    int top[16], left[16], corner;
    for (int i = 0; i < 16; i++) { top[i] = aT ? r[-(ptrdiff_t)f->ys + i] : 0; left[i] = aL ? r[(ptrdiff_t)i * f->ys - 1] : 0; }
    if (!aT && !aL) { for (int i = 0; i < 16; i++) top[i] = left[i] = 128; }
    else if (!aT) { for (int i = 0; i < 16; i++) top[i] = left[0]; }
    else if (!aL) { for (int i = 0; i < 16; i++) left[i] = top[0]; }
    corner = aT && aL ? r[-(ptrdiff_t)f->ys - 1] : aT ? top[0] : aL ? left[0] : 128;
    (void)corner;
    int sum = 16;
    for (int i = 0; i < 16; i++) sum += top[i] + left[i];
    const int dc = sum >> 5;
    int sad[4] = {0, 0, 0, 0};                         /* planar, DC, horizontal, vertical */
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++) {
            const int s = c[(size_t)y * e->cur.ys + x];
            const int pl = ((15 - x) * left[y] + (x + 1) * top[15] + (15 - y) * top[x] + (y + 1) * left[15] + 16) >> 5;
            sad[0] += abs(s - pl); sad[1] += abs(s - dc); sad[2] += abs(s - left[y]); sad[3] += abs(s - top[x]);
        }
    static const int modes[4] = {0, 1, 10, 26};
    int list[3];
    hevc_mpm_list(h, cx, cy, list);
    const int lam = vcp_lambda(h->qp);
    int best = 1;
    unsigned bestc = 0xffffffffu;
    for (int k = 0; k < 4; k++) {
        const unsigned cost = (unsigned)(sad[k] + lam * hevc_mode_bins(list, modes[k]));
        if (cost < bestc) { bestc = cost; best = modes[k]; }
    }
    return best;
}

static void hevc_encode_intra_cu(HEnc* h, int cx, int cy) {
    HCU* cu = &h->cus[cy * h->e->mbw + cx];
    memset(cu, 0, sizeof *cu);
    cu->type = HCU_INTRA;
    cu->imode = 1;
    if (h->e->p.hevc_intra_modes) cu->imode = (uint8_t)hevc_choose_intra_mode(h, cx, cy);
    for (int z = 0; z < 4; z++) {
        const int xl = 16 * cx + 8 * (z & 1), yl = 16 * cy + 8 * (z >> 1);
        uint8_t pred[64];
        hevc_intra_pred(h, 0, xl, yl, 8, cu->imode, pred);
        cu->cbf_y[z] = (uint8_t)hevc_code_block(h, 0, xl, yl, 8, pred, 1, cu->lv_y[z]);
        hevc_intra_pred(h, 1, xl, yl, 4, cu->imode, pred);
        cu->cbf_cb[z] = (uint8_t)hevc_code_block(h, 1, xl >> 1, yl >> 1, 4, pred, 1, cu->lv_cb[z]);
        hevc_intra_pred(h, 2, xl, yl, 4, cu->imode, pred);
        cu->cbf_cr[z] = (uint8_t)hevc_code_block(h, 2, xl >> 1, yl >> 1, 4, pred, 1, cu->lv_cr[z]);
    }
}

/* luma sample interpolation (8.5.3.3.3.1), 8-bit: shift1 = 0, shift2 = 6, then the default weighted prediction
 * (x + 32) >> 6.  mv in quarter samples (fractions 1 and 3 with hevc_subpel >= 2). */
static const int hevc_lf[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
static int hevc_luma_pred(const uint8_t* ref, int rs, int x, int y, int mvx, int mvy) {
    const int fx = mvx & 3, fy = mvy & 3;
    const uint8_t* p = ref + (ptrdiff_t)(y + (mvy >> 2)) * rs + x + (mvx >> 2);
    int v;
    if (!fx && !fy) return p[0];
    if (!fy) { v = 0; for (int i = 0; i < 8; i++) v += hevc_lf[fx][i] * p[i - 3]; }
    else if (!fx) { v = 0; for (int i = 0; i < 8; i++) v += hevc_lf[fy][i] * p[(ptrdiff_t)(i - 3) * rs]; }
    else {
        v = 0;
        for (int i = 0; i < 8; i++) {
            int t = 0;
            for (int j = 0; j < 8; j++) t += hevc_lf[fx][j] * p[(ptrdiff_t)(i - 3) * rs + j - 3];
            v += hevc_lf[fy][i] * t;
        }
        v >>= 6;
    }
    return vcp_clip255((v + 32) >> 6);
}
/* Search proxy for quarter-sample positions (hevc_subpel = 2): the rounded average of the two nearest samples of the
 * half-sample grid, paired as H.264 pairs them (vcp_luma_interp.cuh: hpel_points) -- what the device's refine reads
 * out of the half-sample planes it already holds.  Only the RANKING of the eight quarter-sample candidates uses it;
 * the prediction that is coded (hevc_encode_inter_cu) is always the exact interpolation above. */
static int hevc_luma_proxy(const uint8_t* ref, int rs, int x, int y, int qx, int qy) {
    int x1, y1, x2, y2;
    const int xa = (qx - 1) >> 1, ya = (qy - 1) >> 1;
    if (!(qx & 1) && !(qy & 1)) { x1 = x2 = qx >> 1; y1 = y2 = qy >> 1; }
    else if (!(qy & 1)) { x1 = xa; x2 = xa + 1; y1 = y2 = qy >> 1; }
    else if (!(qx & 1)) { x1 = x2 = qx >> 1; y1 = ya; y2 = ya + 1; }
    else {
        x1 = (xa & 1) ? xa : xa + 1; x2 = (xa & 1) ? xa + 1 : xa;
        y1 = (ya & 1) ? ya + 1 : ya; y2 = (ya & 1) ? ya : ya + 1;
    }
    const int a = hevc_luma_pred(ref, rs, x, y, 2 * x1, 2 * y1);
    if (x1 == x2 && y1 == y2) return a;
    return (a + hevc_luma_pred(ref, rs, x, y, 2 * x2, 2 * y2) + 1) >> 1;
}
/* chroma sample interpolation (8.5.3.3.3.2): the luma vector in eighths of a chroma sample, 4-tap filters */
static void hevc_mc_chroma(const uint8_t* ref, int rs, int x0, int y0, int mvx, int mvy, uint8_t* dst /* 8x8 */) {
    const int ix = mvx >> 3, iy = mvy >> 3, fx = mvx & 7, fy = mvy & 7;
    const int8_t* f4 = hevc_chroma_filter[fx];
    const int8_t* g4 = hevc_chroma_filter[fy];
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            const uint8_t* p = ref + (ptrdiff_t)(y0 + y + iy) * rs + x0 + x + ix;
            int v;
            if (!fx && !fy) v = p[0] << 6;
            else if (!fy) v = f4[0] * p[-1] + f4[1] * p[0] + f4[2] * p[1] + f4[3] * p[2];
            else if (!fx) v = g4[0] * p[-rs] + g4[1] * p[0] + g4[2] * p[rs] + g4[3] * p[2 * rs];
            else {
                int t[4];
                for (int k = 0; k < 4; k++) { const uint8_t* q = p + (ptrdiff_t)(k - 1) * rs; t[k] = f4[0] * q[-1] + f4[1] * q[0] + f4[2] * q[1] + f4[3] * q[2]; }
                v = (g4[0] * t[0] + g4[1] * t[1] + g4[2] * t[2] + g4[3] * t[3]) >> 6;
            }
            dst[y * 8 + x] = (uint8_t)vcp_clip255((v + 32) >> 6);
        }
}

/* full-sample refine on the reconstructed reference: pre-pass vector and its 8 neighbours, zero, predictor.
 * Returns 1 when the CU should be coded intra instead (same estimate on the original picture and the same rule
 * as the H.264 path: intra_estimate, vcp_intra_wins -- every CU decides in parallel). */
static int hevc_refine_cu(HEnc* h, int cx, int cy, int16_t mv[2]) {
    Enc* e = h->e;
    const int lam = vcp_lambda(h->qp);
    int pmx, pmy;
    pmv_estimate(e, cx, cy, &pmx, &pmy);
    const int i = cy * e->mbw + cx, px = 16 * cx, py = 16 * cy;
    const uint8_t* c = e->cur.y + (size_t)py * e->cur.ys + px;
    const int fx = e->mvfp[2 * i], fy = e->mvfp[2 * i + 1];
    int cand[11][2], n = 0;
    cand[n][0] = fx; cand[n][1] = fy; n++;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) { if (!dx && !dy) continue; cand[n][0] = fx + dx; cand[n][1] = fy + dy; n++; }
    cand[n][0] = 0; cand[n][1] = 0; n++;
    cand[n][0] = vcp_clip3(-VCP_MV_FP_MAX, VCP_MV_FP_MAX, (pmx + 2) >> 2);
    cand[n][1] = vcp_clip3(-VCP_MV_FP_MAX, VCP_MV_FP_MAX, (pmy + 2) >> 2); n++;
    uint32_t best = 0xffffffffu;
    for (int k = 0; k < n; k++) {
        const int vx = cand[k][0], vy = cand[k][1];
        const uint8_t* r = h->ref->y + (size_t)(py + vy) * h->ref->ys + px + vx;
        const int cost = sad16(c, e->cur.ys, r, h->ref->ys) + lam * (vcp_se_len(4 * vx - pmx) + vcp_se_len(4 * vy - pmy));
        const uint32_t key = ((uint32_t)cost << 4) | (uint32_t)k;
        if (key < best) best = key;
    }
    int bx = 4 * cand[best & 15][0], by = 4 * cand[best & 15][1];
    uint32_t bcost = best >> 4;
    mv[0] = (int16_t)bx; mv[1] = (int16_t)by;
    if (bcost < VCP_SUBPEL_SKIP_COST) return 0;
    /* half-sample neighbours of the best full-sample position, then (hevc_subpel >= 2) the quarter-sample
     * neighbours of that: raster order, strict improvement (the same rule as the H.264 refine).
     * Each position of each reference picture has ONE interpolated value, so the device takes the half-sample
     * ones from three planes built once per picture (k2_hpel.cu).  The quarter-sample candidates are ranked by
     * the proxy above (hevc_subpel = 2: the medium tiers) or by their exact prediction (3: the slow tiers). */
    for (int step = 2; step >= 1 && e->p.hevc_subpel; step--) {
        if (step == 1 && e->p.hevc_subpel < 2) break;
        uint32_t sb = (bcost << 4) | 0;
        int k = 1;
        const int ox = bx, oy = by;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                if (!dx && !dy) continue;
                const int vx = ox + step * dx, vy = oy + step * dy;
                int sad = 0;
                for (int y = 0; y < 16; y++)
                    for (int x = 0; x < 16; x++)
                        sad += abs(c[(size_t)y * e->cur.ys + x] - (step == 1 && e->p.hevc_subpel == 2 ? hevc_luma_proxy(h->ref->y, h->ref->ys, px + x, py + y, vx, vy)
                                                                                                       : hevc_luma_pred(h->ref->y, h->ref->ys, px + x, py + y, vx, vy)));
                const int cost = sad + lam * (vcp_se_len(vx - pmx) + vcp_se_len(vy - pmy));
                const uint32_t key = ((uint32_t)cost << 4) | (uint32_t)k;
                if (key < sb) { sb = key; bx = vx; by = vy; }
                k++;
            }
        bcost = sb >> 4;
        mv[0] = (int16_t)bx; mv[1] = (int16_t)by;
    }
    return vcp_intra_wins(intra_estimate(e, cx, cy), (int)bcost, lam) != 0;
}

static void hevc_encode_inter_cu(HEnc* h, int cx, int cy) {
    Enc* e = h->e;
    HCU* cu = &h->cus[cy * e->mbw + cx];
    const int16_t mvx = cu->mv[0], mvy = cu->mv[1];
    memset(cu, 0, sizeof *cu);
    cu->type = HCU_INTER; cu->mv[0] = mvx; cu->mv[1] = mvy;
    uint8_t pu[64], pv[64];
    hevc_mc_chroma(h->ref->u, h->ref->cs, 8 * cx, 8 * cy, mvx, mvy, pu);
    hevc_mc_chroma(h->ref->v, h->ref->cs, 8 * cx, 8 * cy, mvx, mvy, pv);
    for (int z = 0; z < 4; z++) {
        const int xl = 16 * cx + 8 * (z & 1), yl = 16 * cy + 8 * (z >> 1);
        uint8_t pred[64];
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) pred[y * 8 + x] = (uint8_t)hevc_luma_pred(h->ref->y, h->ref->ys, xl + x, yl + y, mvx, mvy);
        cu->cbf_y[z] = (uint8_t)hevc_code_block(h, 0, xl, yl, 8, pred, 0, cu->lv_y[z]);
        uint8_t pc[16];
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) pc[y * 4 + x] = pu[(4 * (z >> 1) + y) * 8 + 4 * (z & 1) + x];
        cu->cbf_cb[z] = (uint8_t)hevc_code_block(h, 1, xl >> 1, yl >> 1, 4, pc, 0, cu->lv_cb[z]);
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) pc[y * 4 + x] = pv[(4 * (z >> 1) + y) * 8 + 4 * (z & 1) + x];
        cu->cbf_cr[z] = (uint8_t)hevc_code_block(h, 2, xl >> 1, yl >> 1, 4, pc, 0, cu->lv_cr[z]);
    }
}

/* neighbouring CU (whole 16x16 CTBs in raster order): NULL when outside the picture / slice or not yet decoded */
static const HCU* hevc_nb(const HEnc* h, int cx, int cy, int dx, int dy) {
    const Enc* e = h->e;
    const int nx = cx + dx, ny = cy + dy;
    if (nx < 0 || nx >= e->mbw || ny < 0 || ny >= e->mbh) return NULL;
    if (ny < slice_first_row(e, slice_of_row(e, cy))) return NULL;
    if (ny > cy || (ny == cy && nx >= cx)) return NULL;
    return &h->cus[ny * e->mbw + nx];
}
static int hevc_is_inter(const HCU* c) { return c && c->type != HCU_INTRA; }

/* merge candidate 0 (8.5.3.2.2..4 with MaxNumMergeCand = 1): first of A1, B1, B0, (A0 never decoded), B2; else zero */
static void hevc_merge_cand(const HEnc* h, int cx, int cy, int mv[2]) {
    const HCU* c;
    mv[0] = mv[1] = 0;
    if (hevc_is_inter(c = hevc_nb(h, cx, cy, -1, 0))) { mv[0] = c->mv[0]; mv[1] = c->mv[1]; return; }
    if (hevc_is_inter(c = hevc_nb(h, cx, cy, 0, -1))) { mv[0] = c->mv[0]; mv[1] = c->mv[1]; return; }
    if (hevc_is_inter(c = hevc_nb(h, cx, cy, 1, -1))) { mv[0] = c->mv[0]; mv[1] = c->mv[1]; return; }
    if (hevc_is_inter(c = hevc_nb(h, cx, cy, -1, -1))) { mv[0] = c->mv[0]; mv[1] = c->mv[1]; return; }
}
/* AMVP list (8.5.3.2.6/7), one reference picture, temporal candidate disabled */
static void hevc_amvp(const HEnc* h, int cx, int cy, int list[2][2]) {
    const HCU* a1 = hevc_nb(h, cx, cy, -1, 0);
    const HCU *b0 = hevc_nb(h, cx, cy, 1, -1), *b1 = hevc_nb(h, cx, cy, 0, -1), *b2 = hevc_nb(h, cx, cy, -1, -1);
    int have_a = 0, have_b = 0, a[2] = {0, 0}, b[2] = {0, 0};
    if (hevc_is_inter(a1)) { have_a = 1; a[0] = a1->mv[0]; a[1] = a1->mv[1]; }
    const HCU* bb = hevc_is_inter(b0) ? b0 : hevc_is_inter(b1) ? b1 : hevc_is_inter(b2) ? b2 : NULL;
    if (bb) { have_b = 1; b[0] = bb->mv[0]; b[1] = bb->mv[1]; }
    const int is_scaled = a1 != NULL;    /* A0 is never available in raster CTB order; availableA1 as a block */
    if (!is_scaled && have_b) { have_a = 1; a[0] = b[0]; a[1] = b[1]; }   /* B takes A's place; B is then re-derived to the same vector */
    int n = 0;
    if (have_a) { list[n][0] = a[0]; list[n][1] = a[1]; n++; }
    if (have_b && !(have_a && a[0] == b[0] && a[1] == b[1])) { list[n][0] = b[0]; list[n][1] = b[1]; n++; }
    while (n < 2) { list[n][0] = list[n][1] = 0; n++; }
}

/* ---- in-loop deblocking (8.7.2) ------------------------------------------------------------------
 * Edges lie on the 8x8 luma grid.  All vertical edges of the picture are filtered first, then all horizontal
 * edges on the result; edges of one direction never overlap (a filter reads 4 and changes at most 3 samples
 * on each side), so both passes are order-free -- the reason the GPU path needs no wavefront here.
 * pps_loop_filter_across_slices_enabled_flag = 0: the top edge of a slice is not filtered. */
static int hevc_bs(const HEnc* h, int xq, int yq, int vertical) {
    const Enc* e = h->e;
    const int xp = vertical ? xq - 1 : xq, yp = vertical ? yq : yq - 1;
    if (xp < 0 || yp < 0) return 0;
    const HCU* cq = &h->cus[(yq >> 4) * e->mbw + (xq >> 4)];
    const HCU* cp = &h->cus[(yp >> 4) * e->mbw + (xp >> 4)];
    if (!vertical && (yq & 15) == 0 && slice_first_row(e, slice_of_row(e, yq >> 4)) == (yq >> 4)) return 0;
    if (cp->type == HCU_INTRA || cq->type == HCU_INTRA) return 2;
    const int zq = ((yq >> 3) & 1) * 2 + ((xq >> 3) & 1), zp = ((yp >> 3) & 1) * 2 + ((xp >> 3) & 1);
    if ((cq->type != HCU_SKIP && cq->cbf_y[zq]) || (cp->type != HCU_SKIP && cp->cbf_y[zp])) return 1;
    if (cp == cq) return 0;
    return abs(cp->mv[0] - cq->mv[0]) >= 4 || abs(cp->mv[1] - cq->mv[1]) >= 4;
}
/* one 4-line segment of a luma edge; p points at q0 of line 0, `step` walks across the edge, `line` along it */
static void hevc_filter_luma(uint8_t* q0, ptrdiff_t step, ptrdiff_t line, int beta, int tc) {
#define P(i, l) ((int)q0[-(ptrdiff_t)((i) + 1) * step + (ptrdiff_t)(l) * line])
#define Q(i, l) ((int)q0[(ptrdiff_t)(i) * step + (ptrdiff_t)(l) * line])
    const int dp0 = abs(P(2, 0) - 2 * P(1, 0) + P(0, 0)), dp3 = abs(P(2, 3) - 2 * P(1, 3) + P(0, 3));
    const int dq0 = abs(Q(2, 0) - 2 * Q(1, 0) + Q(0, 0)), dq3 = abs(Q(2, 3) - 2 * Q(1, 3) + Q(0, 3));
    const int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3;
    if (dpq0 + dpq3 >= beta) return;
    const int s0 = 2 * dpq0 < (beta >> 2) && abs(P(3, 0) - P(0, 0)) + abs(Q(0, 0) - Q(3, 0)) < (beta >> 3) && abs(P(0, 0) - Q(0, 0)) < ((5 * tc + 1) >> 1);
    const int s3 = 2 * dpq3 < (beta >> 2) && abs(P(3, 3) - P(0, 3)) + abs(Q(0, 3) - Q(3, 3)) < (beta >> 3) && abs(P(0, 3) - Q(0, 3)) < ((5 * tc + 1) >> 1);
    const int side = (beta + (beta >> 1)) >> 3, dep = dp < side, deq = dq < side;
    for (int l = 0; l < 4; l++) {
        const int p0 = P(0, l), p1 = P(1, l), p2 = P(2, l), p3 = P(3, l), a0 = Q(0, l), a1 = Q(1, l), a2 = Q(2, l), a3 = Q(3, l);
        uint8_t* c = q0 + (ptrdiff_t)l * line;
        if (s0 && s3) {
            c[-1 * step] = (uint8_t)vcp_clip3(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * a0 + a1 + 4) >> 3);
            c[-2 * step] = (uint8_t)vcp_clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + a0 + 2) >> 2);
            c[-3 * step] = (uint8_t)vcp_clip3(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + a0 + 4) >> 3);
            c[0] = (uint8_t)vcp_clip3(a0 - 2 * tc, a0 + 2 * tc, (p1 + 2 * p0 + 2 * a0 + 2 * a1 + a2 + 4) >> 3);
            c[step] = (uint8_t)vcp_clip3(a1 - 2 * tc, a1 + 2 * tc, (p0 + a0 + a1 + a2 + 2) >> 2);
            c[2 * step] = (uint8_t)vcp_clip3(a2 - 2 * tc, a2 + 2 * tc, (p0 + a0 + a1 + 3 * a2 + 2 * a3 + 4) >> 3);
        } else {
            int d = (9 * (a0 - p0) - 3 * (a1 - p1) + 8) >> 4;
            if (abs(d) >= tc * 10) continue;
            d = vcp_clip3(-tc, tc, d);
            c[-1 * step] = (uint8_t)vcp_clip255(p0 + d);
            c[0] = (uint8_t)vcp_clip255(a0 - d);
            if (dep) c[-2 * step] = (uint8_t)vcp_clip255(p1 + vcp_clip3(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + d) >> 1));
            if (deq) c[step] = (uint8_t)vcp_clip255(a1 + vcp_clip3(-(tc >> 1), tc >> 1, (((a2 + a0 + 1) >> 1) - a1 - d) >> 1));
        }
    }
#undef P
#undef Q
}
static void hevc_filter_chroma(uint8_t* q0, ptrdiff_t step, ptrdiff_t line, int tc) {
    for (int l = 0; l < 2; l++) {      /* 4 luma lines = 2 chroma lines */
        uint8_t* c = q0 + (ptrdiff_t)l * line;
        const int p0 = c[-step], p1 = c[-2 * step], a0 = c[0], a1 = c[step];
        const int d = vcp_clip3(-tc, tc, (((a0 - p0) << 2) + p1 - a1 + 4) >> 3);
        c[-step] = (uint8_t)vcp_clip255(p0 + d);
        c[0] = (uint8_t)vcp_clip255(a0 - d);
    }
}
static void hevc_deblock_picture(HEnc* h) {
    const Enc* e = h->e;
    Frame* f = h->rec;
    for (int vertical = 1; vertical >= 0; vertical--)
        for (int y = 0; y < e->ch; y += vertical ? 4 : 8)
            for (int x = 0; x < e->cw; x += vertical ? 8 : 4) {
                const int bs = hevc_bs(h, x, y, vertical);
                if (!bs) continue;
                const int beta = hevc_beta_tab[vcp_clip3(0, 51, h->qp)], tc = hevc_tc_tab[vcp_clip3(0, 53, h->qp + 2 * (bs - 1))];
                hevc_filter_luma(f->y + (size_t)y * f->ys + x, vertical ? 1 : f->ys, vertical ? f->ys : 1, beta, tc);
                if (bs == 2 && ((vertical ? x : y) & 15) == 0) {
                    const int tcc = hevc_tc_tab[vcp_clip3(0, 53, h->qpc + 2)];
                    hevc_filter_chroma(f->u + (size_t)(y >> 1) * f->cs + (x >> 1), vertical ? 1 : f->cs, vertical ? f->cs : 1, tcc);
                    hevc_filter_chroma(f->v + (size_t)(y >> 1) * f->cs + (x >> 1), vertical ? 1 : f->cs, vertical ? f->cs : 1, tcc);
                }
            }
}

/* ---- sample adaptive offset (8.7.3), luma edge offsets only ------------------------------------------------
 * Decision and application both work on the DEBLOCKED picture; a sample whose neighbour (along the class
 * direction) lies outside the picture or in another slice keeps its value (slices are never filtered across).
 * Per CTB: statistics of (source - deblocked) per class and category, offsets = clipped rounded means, the class
 * with the lowest distortion + rate cost if it beats "off" (vcp_algo.h).  CTB-parallel on the device. */
static const int hevc_sao_dx[4][2] = {{-1, 1}, {0, 0}, {-1, 1}, {1, -1}}, hevc_sao_dy[4][2] = {{0, 0}, {-1, 1}, {-1, 1}, {-1, 1}};
static int hevc_sao_category(const HEnc* h, const uint8_t* d, int ds, int x, int y, int cls) {
    const Enc* e = h->e;
    int sgn = 0;
    for (int k = 0; k < 2; k++) {
        const int nx = x + hevc_sao_dx[cls][k], ny = y + hevc_sao_dy[cls][k];
        if (nx < 0 || ny < 0 || nx >= e->cw || ny >= e->ch) return 0;
        if (slice_of_row(e, ny >> 4) != slice_of_row(e, y >> 4)) return 0;
        const int c = d[(size_t)y * ds + x], n = d[(size_t)ny * ds + nx];
        sgn += c < n ? -1 : c > n ? 1 : 0;
    }
    return sgn == -2 ? 1 : sgn == -1 ? 2 : sgn == 1 ? 3 : sgn == 2 ? 4 : 0;
}
static void hevc_sao_picture(HEnc* h) {
    const Enc* e = h->e;
    Frame* f = h->rec;
    const int lam = vcp_lambda(h->qp), lam2 = lam * lam;
    for (int y = 0; y < e->ch; y++) memcpy(h->sao_tmp + (size_t)y * e->cw, f->y + (size_t)y * f->ys, (size_t)e->cw);
    const uint8_t* d = h->sao_tmp;
    for (int cy = 0; cy < e->mbh; cy++)
        for (int cx = 0; cx < e->mbw; cx++) {
            const int i = cy * e->mbw + cx;
            long long best = lam2;                 /* "off": one bin */
            h->sao_type[i] = 0;
            for (int cls = 0; cls < 4; cls++) {
                int sum[4] = {0, 0, 0, 0}, cnt[4] = {0, 0, 0, 0}, off[4];
                for (int y = 16 * cy; y < 16 * cy + 16; y++)
                    for (int x = 16 * cx; x < 16 * cx + 16; x++) {
                        const int k = hevc_sao_category(h, d, e->cw, x, y, cls);
                        if (k) { sum[k - 1] += e->cur.y[(size_t)y * e->cur.ys + x] - d[(size_t)y * e->cw + x]; cnt[k - 1]++; }
                    }
                const long long c = vcp_sao_class_cost(sum, cnt, lam2, off);
                if (c < best && (off[0] | off[1] | off[2] | off[3])) {
                    best = c; h->sao_type[i] = (uint8_t)(1 + cls);
                    for (int k = 0; k < 4; k++) h->sao_off[4 * i + k] = (int8_t)off[k];
                }
            }
            if (!h->sao_type[i]) continue;
            for (int y = 16 * cy; y < 16 * cy + 16; y++)
                for (int x = 16 * cx; x < 16 * cx + 16; x++) {
                    const int k = hevc_sao_category(h, d, e->cw, x, y, h->sao_type[i] - 1);
                    if (k) f->y[(size_t)y * f->ys + x] = (uint8_t)vcp_clip255(d[(size_t)y * e->cw + x] + h->sao_off[4 * i + k - 1]);
                }
        }
}
/* sao() syntax of one CTU (7.3.8.3): no merging, luma only */
static void hevc_write_sao(const HEnc* h, Cabac* c, int cx, int cy, int row0) {
    const int i = cy * h->e->mbw + cx, type = h->sao_type[i];
    if (cx > 0) cabac_encode(c, HC_SAO_MERGE, 0);
    if (cy > row0) cabac_encode(c, HC_SAO_MERGE, 0);
    cabac_encode(c, HC_SAO_TYPE, type != 0);
    if (!type) return;
    cabac_bypass(c, 1);                            /* sao_type_idx_luma = 2: edge offset */
    for (int k = 0; k < 4; k++) {
        const int a = abs(h->sao_off[4 * i + k]);
        for (int j = 0; j < a; j++) cabac_bypass(c, 1);
        if (a < 7) cabac_bypass(c, 0);
    }
    cabac_bypass(c, ((type - 1) >> 1) & 1); cabac_bypass(c, (type - 1) & 1);    /* sao_eo_class_luma */
}

/* ---- syntax ------------------------------------------------------------------------------------- */
static void hevc_write_tu_tree(Cabac* c, const HCU* cu, int intra) {
    int any_cb = 0, any_cr = 0;
    for (int z = 0; z < 4; z++) { any_cb |= cu->cbf_cb[z]; any_cr |= cu->cbf_cr[z]; }
    /* trafoDepth 0 (16x16, split inferred): chroma cbf */
    cabac_encode(c, HC_CBF_CHROMA + 0, any_cb);
    cabac_encode(c, HC_CBF_CHROMA + 0, any_cr);
    for (int z = 0; z < 4; z++) {
        /* trafoDepth 1 (8x8): chroma cbf if the parent had one, luma cbf always */
        if (any_cb) cabac_encode(c, HC_CBF_CHROMA + 1, cu->cbf_cb[z]);
        if (any_cr) cabac_encode(c, HC_CBF_CHROMA + 1, cu->cbf_cr[z]);
        cabac_encode(c, HC_CBF_LUMA + 0, cu->cbf_y[z]);
        const int scan = hevc_scan_idx(cu->type == HCU_INTRA, cu->imode);   /* luma 8x8 and chroma 4x4 both qualify */
        if (cu->cbf_y[z]) hevc_residual(c, cu->lv_y[z], 3, 0, scan);
        if (cu->cbf_cb[z]) hevc_residual(c, cu->lv_cb[z], 2, 1, scan);
        if (cu->cbf_cr[z]) hevc_residual(c, cu->lv_cr[z], 2, 2, scan);
    }
    (void)intra;
}
static void hevc_write_mvd(Cabac* c, int dx, int dy) {
    const int ax = abs(dx), ay = abs(dy);
    cabac_encode(c, HC_MVD_GT0, ax > 0);
    cabac_encode(c, HC_MVD_GT0, ay > 0);
    if (ax) cabac_encode(c, HC_MVD_GT1, ax > 1);
    if (ay) cabac_encode(c, HC_MVD_GT1, ay > 1);
    if (ax) { if (ax > 1) cabac_ueg_bypass(c, (unsigned)(ax - 2), 1); cabac_bypass(c, dx < 0); }
    if (ay) { if (ay > 1) cabac_ueg_bypass(c, (unsigned)(ay - 2), 1); cabac_bypass(c, dy < 0); }
}
static unsigned long long hevc_write_slice_data(HEnc* h, BW* b, int r0, int r1) {
    Enc* e = h->e;
    Cabac c; hevc_cabac_init(&c, b, h->idr ? 0 : 1, h->qp);
    for (int cy = r0; cy < r1; cy++)
        for (int cx = 0; cx < e->mbw; cx++) {
            const HCU* cu = &h->cus[cy * e->mbw + cx];
            if (e->p.hevc_sao) hevc_write_sao(h, &c, cx, cy, r0);
            if (!h->idr) {
                const HCU *l = hevc_nb(h, cx, cy, -1, 0), *a = hevc_nb(h, cx, cy, 0, -1);
                cabac_encode(&c, HC_SKIP + (l && l->type == HCU_SKIP) + (a && a->type == HCU_SKIP), cu->type == HCU_SKIP);
            }
            if (cu->type != HCU_SKIP) {
                if (!h->idr) cabac_encode(&c, HC_PRED_MODE, cu->type == HCU_INTRA);
                cabac_encode(&c, HC_PART_MODE, 1);                       /* PART_2Nx2N */
                if (cu->type == HCU_INTRA) {
                    int list[3];
                    hevc_mpm_list(h, cx, cy, list);
                    const int mi = cu->imode == list[0] ? 0 : cu->imode == list[1] ? 1 : cu->imode == list[2] ? 2 : -1;
                    cabac_encode(&c, HC_PREV_INTRA, mi >= 0);              /* prev_intra_luma_pred_flag */
                    if (mi >= 0) { cabac_bypass(&c, mi > 0); if (mi > 0) cabac_bypass(&c, mi > 1); }   /* mpm_idx: 0, 10, 11 */
                    else {                                                 /* rem_intra_luma_pred_mode: 5 bits */
                        int srt[3] = {list[0], list[1], list[2]}, rem = cu->imode;
                        for (int a = 0; a < 2; a++) for (int b2 = a + 1; b2 < 3; b2++) if (srt[a] > srt[b2]) { const int t2 = srt[a]; srt[a] = srt[b2]; srt[b2] = t2; }
                        for (int a = 2; a >= 0; a--) if (rem > srt[a]) rem--;
                        for (int q = 4; q >= 0; q--) cabac_bypass(&c, (rem >> q) & 1);
                    }
                    cabac_encode(&c, HC_CHROMA_MODE, 0);                   /* intra_chroma_pred_mode 4: derived from luma */
                    hevc_write_tu_tree(&c, cu, 1);
                } else {
                    cabac_encode(&c, HC_MERGE_FLAG, cu->merge);
                    if (!cu->merge) {
                        hevc_write_mvd(&c, cu->mvd[0], cu->mvd[1]);
                        cabac_encode(&c, HC_MVP_FLAG, cu->mvp_idx);
                    }
                    int any = 0;
                    for (int z = 0; z < 4; z++) any |= cu->cbf_y[z] | cu->cbf_cb[z] | cu->cbf_cr[z];
                    /* rqt_root_cbf is not sent for a 2Nx2N merge CU (inferred 1: without residual it would be a skip) */
                    if (!cu->merge) cabac_encode(&c, HC_RQT_ROOT_CBF, any);
                    if (any) hevc_write_tu_tree(&c, cu, 0);
                }
            }
            cabac_terminate(&c, cy == r1 - 1 && cx == e->mbw - 1);       /* end_of_slice_segment_flag */
        }
    while (b->nbits) bw_put(b, 1, 0);
    return c.nbins;
}

/* ---- parameter sets and slice header ----------------------------------------------------------- */
static size_t hevc_nal_write(uint8_t* out, size_t cap, int type, const uint8_t* rbsp, size_t n) {
    size_t o = 0; int zeros = 0;
    if (cap < 6) return 0;
    out[o++] = 0; out[o++] = 0; out[o++] = 0; out[o++] = 1;
    out[o++] = (uint8_t)(type << 1); out[o++] = 1;    /* nuh_layer_id 0, nuh_temporal_id_plus1 1 */
    for (size_t i = 0; i < n; i++) {
        if (zeros >= 2 && rbsp[i] <= 3) { if (o >= cap) return 0; out[o++] = 3; zeros = 0; }
        if (o >= cap) return 0;
        out[o++] = rbsp[i];
        zeros = rbsp[i] == 0 ? zeros + 1 : 0;
    }
    return o;
}
static int hevc_level_idc(const Enc* e) {
    /* table A.8: MaxLumaPs / MaxLumaSr -> general_level_idc = 30 * level */
    static const struct { int idc; long ps; double sr; } L[] = {
        {30, 36864, 552960}, {60, 122880, 3686400}, {63, 245760, 7372800}, {90, 552960, 16588800}, {93, 983040, 33177600},
        {120, 2228224, 66846720}, {123, 2228224, 133693440}, {150, 8912896, 267386880}, {153, 8912896, 534773760},
        {156, 8912896, 1069547520}, {180, 35651584, 1069547520}, {183, 35651584, 2139095040.0}, {186, 35651584, 4278190080.0}};
    const long ps = (long)e->cw * e->ch;
    const double sr = (double)ps * e->p.fps_num / (e->p.fps_den > 0 ? e->p.fps_den : 1);
    for (unsigned i = 0; i < sizeof L / sizeof L[0]; i++) if (ps <= L[i].ps && sr <= L[i].sr) return L[i].idc;
    return 186;
}
static void hevc_ptl(BW* b, const Enc* e) {
    bw_put(b, 2, 0); bw_put(b, 1, 0); bw_put(b, 5, 1);          /* profile space 0, main tier, Main */
    bw_put32(b, 0x60000000u);                                   /* compatible with Main and Main 10 */
    bw_put(b, 1, 1); bw_put(b, 1, 0); bw_put(b, 1, 0); bw_put(b, 1, 1);   /* progressive, !interlaced, !non-packed, frame-only */
    bw_put(b, 22, 0); bw_put(b, 22, 0);                         /* 43 reserved bits + inbld */
    bw_put(b, 8, (uint32_t)hevc_level_idc(e));
}
static size_t hevc_write_vps(const Enc* e, uint8_t* out, size_t cap) {
    uint8_t tmp[128]; BW b; bw_init(&b, tmp, sizeof tmp);
    bw_put(&b, 4, 0); bw_put(&b, 1, 1); bw_put(&b, 1, 1); bw_put(&b, 6, 0); bw_put(&b, 3, 0); bw_put(&b, 1, 1);
    bw_put(&b, 16, 0xffff);
    hevc_ptl(&b, e);
    bw_put(&b, 1, 1);                         /* vps_sub_layer_ordering_info_present_flag */
    bw_ue(&b, 1); bw_ue(&b, 0); bw_ue(&b, 0); /* max_dec_pic_buffering_minus1, max_num_reorder_pics, max_latency_increase_plus1 */
    bw_put(&b, 6, 0); bw_ue(&b, 0);           /* vps_max_layer_id, vps_num_layer_sets_minus1 */
    bw_put(&b, 1, 0);                         /* vps_timing_info_present_flag */
    bw_put(&b, 1, 0);                         /* vps_extension_flag */
    bw_trailing(&b);
    return hevc_nal_write(out, cap, 32, tmp, b.pos);
}
static size_t hevc_write_sps(const Enc* e, uint8_t* out, size_t cap) {
    uint8_t tmp[160]; BW b; bw_init(&b, tmp, sizeof tmp);
    const vcpenc_params* p = &e->p;
    bw_put(&b, 4, 0); bw_put(&b, 3, 0); bw_put(&b, 1, 1);
    hevc_ptl(&b, e);
    bw_ue(&b, 0);                             /* sps_seq_parameter_set_id */
    bw_ue(&b, 1);                             /* chroma_format_idc 4:2:0 */
    bw_ue(&b, (unsigned)e->cw); bw_ue(&b, (unsigned)e->ch);
    const int cr = e->cw - p->width, cbm = e->ch - p->height;
    if (cr || cbm) { bw_put(&b, 1, 1); bw_ue(&b, 0); bw_ue(&b, (unsigned)(cr / 2)); bw_ue(&b, 0); bw_ue(&b, (unsigned)(cbm / 2)); }
    else bw_put(&b, 1, 0);
    bw_ue(&b, 0); bw_ue(&b, 0);               /* bit depths */
    bw_ue(&b, 4);                             /* log2_max_pic_order_cnt_lsb_minus4 -> 8 bits */
    bw_put(&b, 1, 1); bw_ue(&b, 1); bw_ue(&b, 0); bw_ue(&b, 0);   /* sub-layer ordering info */
    bw_ue(&b, 1);                             /* log2_min_luma_coding_block_size_minus3: 16 */
    bw_ue(&b, 0);                             /* log2_diff_max_min_luma_coding_block_size: CTB 16 */
    bw_ue(&b, 0);                             /* log2_min_luma_transform_block_size_minus2: 4 */
    bw_ue(&b, 1);                             /* log2_diff_max_min_luma_transform_block_size: 8 */
    bw_ue(&b, 0); bw_ue(&b, 0);               /* max_transform_hierarchy_depth_inter / intra */
    bw_put(&b, 1, 0);                         /* scaling_list_enabled_flag */
    bw_put(&b, 1, 0);                         /* amp_enabled_flag */
    bw_put(&b, 1, p->hevc_sao ? 1 : 0);       /* sample_adaptive_offset_enabled_flag */
    bw_put(&b, 1, 0);                         /* pcm_enabled_flag */
    bw_ue(&b, 1);                             /* num_short_term_ref_pic_sets */
    bw_ue(&b, 1); bw_ue(&b, 0); bw_ue(&b, 0); bw_put(&b, 1, 1);   /* one negative picture, delta_poc -1, used */
    bw_put(&b, 1, 0);                         /* long_term_ref_pics_present_flag */
    bw_put(&b, 1, 0);                         /* sps_temporal_mvp_enabled_flag */
    bw_put(&b, 1, 0);                         /* strong_intra_smoothing_enabled_flag */
    bw_put(&b, 1, 1);                         /* vui_parameters_present_flag */
    {
        bw_put(&b, 1, 0); bw_put(&b, 1, 0); bw_put(&b, 1, 0); bw_put(&b, 1, 0);   /* aspect, overscan, video signal, chroma loc */
        bw_put(&b, 1, 0); bw_put(&b, 1, 0); bw_put(&b, 1, 0);                     /* neutral chroma, field seq, frame field info */
        bw_put(&b, 1, 0);                     /* default_display_window_flag */
        bw_put(&b, 1, 1);                     /* vui_timing_info_present_flag */
        bw_put32(&b, (uint32_t)p->fps_den); bw_put32(&b, (uint32_t)p->fps_num);
        bw_put(&b, 1, 0);                     /* vui_poc_proportional_to_timing_flag */
        bw_put(&b, 1, 0);                     /* vui_hrd_parameters_present_flag */
        bw_put(&b, 1, 0);                     /* bitstream_restriction_flag */
    }
    bw_put(&b, 1, 0);                         /* sps_extension_present_flag */
    bw_trailing(&b);
    return hevc_nal_write(out, cap, 33, tmp, b.pos);
}
static size_t hevc_write_pps(const Enc* e, uint8_t* out, size_t cap) {
    uint8_t tmp[64]; BW b; bw_init(&b, tmp, sizeof tmp);
    bw_ue(&b, 0); bw_ue(&b, 0);
    bw_put(&b, 1, 0); bw_put(&b, 1, 0); bw_put(&b, 3, 0);       /* dependent slices, output flag, extra header bits */
    bw_put(&b, 1, 0); bw_put(&b, 1, 0);                         /* sign_data_hiding, cabac_init_present */
    bw_ue(&b, 0); bw_ue(&b, 0);                                 /* num_ref_idx defaults */
    bw_se(&b, 0);                                               /* init_qp_minus26 */
    bw_put(&b, 1, 0); bw_put(&b, 1, 0); bw_put(&b, 1, 0);       /* constrained_intra_pred, transform_skip, cu_qp_delta */
    bw_se(&b, 0); bw_se(&b, 0);                                 /* cb / cr qp offsets */
    bw_put(&b, 1, 0); bw_put(&b, 1, 0); bw_put(&b, 1, 0); bw_put(&b, 1, 0);   /* slice chroma offsets, weighted pred x2, transquant bypass */
    bw_put(&b, 1, 0); bw_put(&b, 1, 0);                         /* tiles, entropy_coding_sync */
    bw_put(&b, 1, 0);                                           /* pps_loop_filter_across_slices_enabled_flag */
    if (e->p.deblock_idc == 1) { bw_put(&b, 1, 1); bw_put(&b, 1, 0); bw_put(&b, 1, 1); }   /* deblocking control present, no override, DISABLED */
    else bw_put(&b, 1, 0);                                      /* no deblocking control: enabled, beta / tc offsets 0 */
    bw_put(&b, 1, 0); bw_put(&b, 1, 0);                         /* scaling list data, lists modification */
    bw_ue(&b, 0);                                               /* log2_parallel_merge_level_minus2 */
    bw_put(&b, 1, 0); bw_put(&b, 1, 0);                         /* slice header extension, pps extension */
    bw_trailing(&b);
    return hevc_nal_write(out, cap, 34, tmp, b.pos);
}
static void hevc_write_slice_header(const Enc* e, BW* b, int first_ctb, int idr, int poc, int qp) {
    bw_put(b, 1, first_ctb == 0);
    if (idr) bw_put(b, 1, 0);                 /* no_output_of_prior_pics_flag */
    bw_ue(b, 0);                              /* slice_pic_parameter_set_id */
    if (first_ctb) {
        int bits = 0;
        while ((1 << bits) < e->nmb) bits++;
        bw_put(b, bits, (uint32_t)first_ctb); /* slice_segment_address */
    }
    bw_ue(b, idr ? 2 : 1);                    /* slice_type: I / P */
    if (!idr) {
        bw_put(b, 8, (uint32_t)(poc & 255));  /* slice_pic_order_cnt_lsb */
        bw_put(b, 1, 1);                      /* short_term_ref_pic_set_sps_flag (one set: no index) */
    }
    if (e->p.hevc_sao) { bw_put(b, 1, 1); bw_put(b, 1, 0); }   /* slice_sao_luma_flag, slice_sao_chroma_flag */
    if (!idr) {
        bw_put(b, 1, 0);                      /* num_ref_idx_active_override_flag */
        bw_ue(b, 4);                          /* five_minus_max_num_merge_cand: 1 candidate */
    }
    bw_se(b, qp - 26);                        /* slice_qp_delta */
    bw_put(b, 1, 1);                          /* byte_alignment(): one, then zeros */
    while (b->nbits) bw_put(b, 1, 0);
}

int hevc_orc_encode(const vcpenc_params* p, const uint8_t* frames, int nframes, uint8_t* out, size_t out_cap,
                    size_t* out_len, vcpenc_frame_info* info, uint8_t* recon) {
    Enc enc; Enc* e = &enc;
    int rc = VCPENC_E_INTERNAL;
    memset(e, 0, sizeof *e);
    e->p = *p;
    if (p->width < 16 || p->height < 16 || (p->width & 1) || (p->height & 1) || p->gop < 1 || p->slices < 0) return VCPENC_E_ARGS;
    if (p->in_fmt != VCPENC_FMT_YUV420P) return VCPENC_E_ARGS;
    if (p->rc_mode == VCPENC_RC_ABR && (p->bitrate <= 0 || p->fps_num <= 0 || p->fps_den <= 0)) return VCPENC_E_ARGS;
    e->mbw = (p->width + 15) / 16; e->mbh = (p->height + 15) / 16; e->nmb = e->mbw * e->mbh;
    if (e->p.slices == 0) e->p.slices = vcp_auto_slices(e->mbh, 1);
    if (e->p.slices > e->mbh) return VCPENC_E_ARGS;
    p = &e->p;
    e->cw = 16 * e->mbw; e->ch = 16 * e->mbh;
    HEnc h; memset(&h, 0, sizeof h); h.e = e;
    if (frame_alloc(&e->cur, e->cw, e->ch) || frame_alloc(&e->prev_orig, e->cw, e->ch) ||
        frame_alloc(&e->recon[0], e->cw, e->ch) || frame_alloc(&e->recon[1], e->cw, e->ch) ||
        half_alloc(&e->hcur, e->cw, e->ch) || half_alloc(&e->hprev, e->cw, e->ch)) goto done;
    h.cus = (HCU*)calloc((size_t)e->nmb, sizeof(HCU));
    h.sao_type = (uint8_t*)calloc((size_t)e->nmb, 1);
    h.sao_off = (int8_t*)calloc((size_t)e->nmb, 4);
    h.sao_tmp = (uint8_t*)malloc((size_t)e->cw * e->ch);
    e->mvfp = (int16_t*)calloc((size_t)e->nmb * 2, sizeof(int16_t));
    e->rbsp_cap = (size_t)e->nmb * 1024 + 4096;
    e->rbsp = (uint8_t*)malloc(e->rbsp_cap);
    if (!h.cus || !e->mvfp || !e->rbsp || !h.sao_type || !h.sao_off || !h.sao_tmp) goto done;
    const size_t fsz = (size_t)p->width * p->height + 2 * (size_t)((p->width + 1) / 2) * ((p->height + 1) / 2);
    size_t o = 0;
    int ri = 0;
    /* rate control: the same model and two-picture feedback delay as the H.264 path (vcp_algo.h) */
    const int abr = p->rc_mode == VCPENC_RC_ABR;
    const int vbv = vcp_rc_has_vbv(p->maxrate, p->bufsize, p->fps_num, p->fps_den);
    const int rc_fb = abr || vbv;
    const int rc_qp0 = abr ? vcp_rc_initial_qp(vcp_rc_eff_bitrate(p->bitrate, p->maxrate), p->fps_num, p->fps_den, p->width, p->height) : 0;
    unsigned long long rc_cum = 0;
    long long rc_full = 0;
    int rc_qp_next[2] = {0, 0};
    for (int n = 0; n < nframes; n++) {
        const int t = n % p->gop, idr = t == 0;
        int qp = idr ? p->qp_i : p->qp_p;
        if (rc_fb) {
            if (idr) {
                rc_qp_next[0] = rc_qp_next[1] = abr ? rc_qp0 : p->qp_p;
                if (abr) { qp = rc_qp0 - VCP_RC_QP_I_OFFSET; if (qp < 0) qp = 0; }
            } else { qp = rc_qp_next[0]; rc_qp_next[0] = rc_qp_next[1]; }
        }
        h.qp = qp; h.qpc = hevc_chroma_qp(qp); h.idr = idr;
        { Frame tf = e->prev_orig; e->prev_orig = e->cur; e->cur = tf; Half th = e->hprev; e->hprev = e->hcur; e->hcur = th; }
        frame_load_yuv420p(&e->cur, frames + (size_t)n * fsz, p->width, p->height);
        half_build(&e->hcur, &e->cur);
        h.rec = &e->recon[ri]; h.ref = &e->recon[ri ^ 1];
        if (idr) {
            for (int cy = 0; cy < e->mbh; cy++) for (int cx = 0; cx < e->mbw; cx++) hevc_encode_intra_cu(&h, cx, cy);
        } else {
            me_prepass(e);
            for (int i = 0; i < e->nmb; i++) {
                int16_t mv[2];
                const int intra = hevc_refine_cu(&h, i % e->mbw, i / e->mbw, mv);
                h.cus[i].mv[0] = mv[0]; h.cus[i].mv[1] = mv[1];
                if (intra) h.cus[i].type = HCU_INTRA;     /* coded below, once the inter CUs around it are reconstructed */
                else hevc_encode_inter_cu(&h, i % e->mbw, i / e->mbw);
            }
            /* intra CUs inside the P picture (scene cuts): raster order, predicting from whatever surrounds them */
            for (int i = 0; i < e->nmb; i++)
                if (h.cus[i].type == HCU_INTRA) hevc_encode_intra_cu(&h, i % e->mbw, i / e->mbw);
            /* post-hoc: merge / skip / AMVP index, once every vector of the picture is final */
            for (int i = 0; i < e->nmb; i++) {
                HCU* cu = &h.cus[i];
                if (cu->type == HCU_INTRA) continue;
                const int cx = i % e->mbw, cy = i / e->mbw;
                int mg[2], list[2][2], any = 0;
                hevc_merge_cand(&h, cx, cy, mg);
                for (int z = 0; z < 4; z++) any |= cu->cbf_y[z] | cu->cbf_cb[z] | cu->cbf_cr[z];
                if (mg[0] == cu->mv[0] && mg[1] == cu->mv[1]) {
                    cu->merge = 1;
                    if (!any) cu->type = HCU_SKIP;
                } else {
                    hevc_amvp(&h, cx, cy, list);
                    const int c0 = vcp_se_len(cu->mv[0] - list[0][0]) + vcp_se_len(cu->mv[1] - list[0][1]);
                    const int c1 = vcp_se_len(cu->mv[0] - list[1][0]) + vcp_se_len(cu->mv[1] - list[1][1]);
                    cu->mvp_idx = (uint8_t)(c1 < c0);
                    cu->mvd[0] = (int16_t)(cu->mv[0] - list[cu->mvp_idx][0]);
                    cu->mvd[1] = (int16_t)(cu->mv[1] - list[cu->mvp_idx][1]);
                }
            }
        }
        /* in-loop filters run before the slice data is written: the SAO parameters are part of it */
        if (p->deblock_idc != 1) hevc_deblock_picture(&h);
        if (p->hevc_sao) hevc_sao_picture(&h);
        const size_t au0 = o;
        if (idr) {
            size_t k = hevc_write_vps(e, out + o, out_cap - o); if (!k) { rc = VCPENC_E_OVERFLOW; goto done; } o += k;
            k = hevc_write_sps(e, out + o, out_cap - o); if (!k) { rc = VCPENC_E_OVERFLOW; goto done; } o += k;
            k = hevc_write_pps(e, out + o, out_cap - o); if (!k) { rc = VCPENC_E_OVERFLOW; goto done; } o += k;
        }
        unsigned long long frame_bits = 0;
        for (int s = 0; s < p->slices; s++) {
            const int r0 = slice_first_row(e, s), r1 = s + 1 < p->slices ? slice_first_row(e, s + 1) : e->mbh;
            BW b; bw_init(&b, e->rbsp, e->rbsp_cap);
            hevc_write_slice_header(e, &b, r0 * e->mbw, idr, t, qp);
            frame_bits += (hevc_write_slice_data(&h, &b, r0, r1) * VCP_CABAC_BITS_PER_BIN_Q4) >> 4;
            if (b.overflow) { rc = VCPENC_E_OVERFLOW; goto done; }
            const size_t k = hevc_nal_write(out + o, out_cap - o, idr ? 19 : 1, e->rbsp, b.pos);   /* IDR_W_RADL / TRAIL_R */
            if (!k) { rc = VCPENC_E_OVERFLOW; goto done; }
            o += k;
        }
        if (rc_fb) {
            int gop_len = p->gop;
            if (n - t + gop_len > nframes) gop_len = nframes - (n - t);
            rc_qp_next[1] = vcp_rc_picture(abr, rc_qp0, p->qp_p, abr ? vcp_rc_gop_budget(p->bitrate, p->maxrate, p->fps_num, p->fps_den, gop_len) : 0,
                                           vbv ? vcp_vbv_rate(p->maxrate, p->fps_num, p->fps_den) : 0, vbv ? p->bufsize : 0,
                                           qp, rc_qp_next[0], idr, frame_bits, t, gop_len, &rc_cum, &rc_full);
        }
        if (info) { info[n].offset = au0; info[n].size = (uint32_t)(o - au0); info[n].is_idr = (uint8_t)idr; info[n].qp = (uint8_t)qp; }
        frame_pad(h.rec);
        if (recon) store_recon(e, h.rec, recon + (size_t)n * fsz);
        ri ^= 1;
    }
    *out_len = o;
    rc = VCPENC_OK;
done:
    frame_free(&e->cur); frame_free(&e->prev_orig); frame_free(&e->recon[0]); frame_free(&e->recon[1]);
    free(e->hcur.buf); free(e->hprev.buf); free(h.cus); free(h.sao_type); free(h.sao_off); free(h.sao_tmp); free(e->mvfp); free(e->rbsp);
    return rc;
}
