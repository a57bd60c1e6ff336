// Normative HEVC (ITU-T H.265) constants shared by the CPU oracle and the CUDA kernels: core transform
// matrices (8.6.4.2), quantiser / level scales (8.6.3), chroma QP mapping (table 8-10), CABAC initValues
// (tables 9-5..9-37, cross-checked against the FFmpeg decoder's tables by tests/test_oracle.py), scans (6.5.3).
// Define VCP_TAB before including to choose the storage class (see h264_tables.h).
#ifndef VCP_HEVC_TABLES_H
#define VCP_HEVC_TABLES_H
#include <stdint.h>
#ifndef VCP_TAB
#define VCP_TAB static const
#endif

VCP_TAB int8_t hevc_dct8[8][8] = {
    {64, 64, 64, 64, 64, 64, 64, 64},    {89, 75, 50, 18, -18, -50, -75, -89}, {83, 36, -36, -83, -83, -36, 36, 83},
    {75, -18, -89, -50, 50, 89, 18, -75}, {64, -64, -64, 64, 64, -64, -64, 64}, {50, -89, 18, 75, -75, -18, 89, -50},
    {36, -83, 83, -36, -36, 83, -83, 36}, {18, -50, 75, -89, 89, -75, 50, -18}};
VCP_TAB int8_t hevc_dct4[4][4] = {{64, 64, 64, 64}, {83, 36, -36, -83}, {64, -64, -64, 64}, {36, -83, 83, -36}};
VCP_TAB int hevc_quant_scale[6] = {26214, 23302, 20560, 18396, 16384, 14564};
VCP_TAB int hevc_level_scale[6] = {40, 45, 51, 57, 64, 72};
VCP_TAB uint8_t hevc_qpc_tab[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};   /* qPi 30..43 */

/* context layout of this encoder (initValue per initType 0 = I, 1 = P with cabac_init_flag 0) */
enum {
    HC_SKIP = 0,            /* 3 */
    HC_PRED_MODE = 3,       /* 1 */
    HC_PART_MODE = 4,       /* 1 (first bin only: 2Nx2N) */
    HC_PREV_INTRA = 5,      /* 1 */
    HC_CHROMA_MODE = 6,     /* 1 */
    HC_MERGE_FLAG = 7,      /* 1 */
    HC_MVP_FLAG = 8,        /* 1 */
    HC_MVD_GT0 = 9,         /* 1 */
    HC_MVD_GT1 = 10,        /* 1 */
    HC_RQT_ROOT_CBF = 11,   /* 1 */
    HC_CBF_LUMA = 12,       /* 2 */
    HC_CBF_CHROMA = 14,     /* 4 */
    HC_LAST_X = 18,         /* 18 */
    HC_LAST_Y = 36,         /* 18 */
    HC_CSBF = 54,           /* 4 */
    HC_SIG = 58,            /* 42 */
    HC_GT1 = 100,           /* 24 */
    HC_GT2 = 124,           /* 6 */
    HC_SAO_MERGE = 130,     /* 1: sao_merge_left_flag / sao_merge_up_flag */
    HC_SAO_TYPE = 131,      /* 1: first bin of sao_type_idx_luma / _chroma */
    HC_NCTX = 132
};
VCP_TAB uint8_t hevc_init_values[2][HC_NCTX] = {
    {   /* initType 0 */
        154, 154, 154, /* cu_skip_flag: unused in I */ 154 /* pred_mode: unused */, 184, 184, 63, 154, 168, 154, 154, 154,
        111, 141, 94, 138, 182, 154,
        110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,
        110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,
        91, 171, 134, 141,
        111, 111, 125, 110, 110, 94, 124, 108, 124, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125,
        107, 125, 141, 179, 153, 125, 140, 139, 182, 182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111,
        140, 92, 137, 138, 140, 152, 138, 139, 153, 74, 149, 92, 139, 107, 122, 152, 140, 179, 166, 182, 140, 227, 122, 197,
        138, 153, 136, 167, 152, 152,
        153, 200},
    {   /* initType 1 */
        197, 185, 201, 149, 154, 154, 152, 110, 168, 140, 198, 79,
        153, 111, 149, 107, 167, 154,
        125, 110, 94, 110, 95, 79, 125, 111, 110, 78, 110, 111, 111, 95, 94, 108, 123, 108,
        125, 110, 94, 110, 95, 79, 125, 111, 110, 78, 110, 111, 111, 95, 94, 108, 123, 108,
        121, 140, 61, 154,
        155, 154, 139, 153, 139, 123, 123, 63, 153, 166, 183, 140, 136, 153, 154, 166, 183, 140, 136, 153, 154,
        166, 183, 140, 136, 153, 154, 170, 153, 123, 123, 107, 121, 107, 121, 167, 151, 183, 140, 151, 183, 140,
        154, 196, 196, 167, 154, 152, 167, 182, 182, 134, 149, 136, 153, 121, 136, 137, 169, 194, 166, 167, 154, 167, 137, 182,
        107, 167, 91, 122, 107, 167,
        153, 185}};

/* 4x4 up-right diagonal scan (6.5.3): scan position -> (x, y) */
VCP_TAB uint8_t hevc_diag4_x[16] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 1, 2, 3, 2, 3, 3};
VCP_TAB uint8_t hevc_diag4_y[16] = {0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1, 3, 2, 3};
VCP_TAB uint8_t hevc_diag2_x[4] = {0, 0, 1, 1}, hevc_diag2_y[4] = {0, 1, 0, 1};
VCP_TAB uint8_t hevc_sig_ctx_map4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};   /* index (yC << 2) + xC */


/* deblocking (8.7.2.5.3, table 8-12): beta' by Q = 0..51, tc' by Q = 0..53 */
VCP_TAB uint8_t hevc_beta_tab[52] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24,
    26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
VCP_TAB uint8_t hevc_tc_tab[54] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3,
    3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};

/* chroma sample interpolation filter (table 8-13): [fraction in eighths][tap] */
VCP_TAB int8_t hevc_chroma_filter[8][4] = {{0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
                                           {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};

#endif  // VCP_HEVC_TABLES_H
