// Sequence / picture parameter sets (H.264 7.3.2.1, 7.3.2.2, E.1.1).
#include "host_bits.h"

namespace vcp {

int level_idc_for(int mbw, int mbh, int fps_num, int fps_den) {
    // Table A-1: level, MaxMBPS, MaxFS
    static const struct { int idc; long mbps; long fs; } L[] = {
        {10, 1485, 99},     {11, 3000, 396},    {12, 6000, 396},     {13, 11880, 396},
        {20, 11880, 396},   {21, 19800, 792},   {22, 20250, 1620},   {30, 40500, 1620},
        {31, 108000, 3600}, {32, 216000, 5120}, {40, 245760, 8192},  {42, 522240, 8704},
        {50, 589824, 22080}, {51, 983040, 36864}, {52, 2073600, 36864}};
    const long fs = (long)mbw * mbh;
    const long mbps = (long)((double)fs * fps_num / (fps_den > 0 ? fps_den : 1) + 0.5);
    for (const auto& l : L)
        if (fs <= l.fs && mbps <= l.mbps) return l.idc;
    return 52;
}

std::vector<uint8_t> make_sps_nal(const vcpenc_params& p) {
    const int mbw = (p.width + 15) / 16, mbh = (p.height + 15) / 16;
    BitWriter b;
    if (p.transform8x8) { b.put(8, 100); b.put(8, 0x00); }   // High
    else if (p.entropy) { b.put(8, 77); b.put(8, 0x40); }    // Main (CABAC), constraint_set1
    else { b.put(8, 66); b.put(8, 0xC0); }                   // Constrained Baseline: constraint_set0/1
    b.put(8, (uint32_t)level_idc_for(mbw, mbh, p.fps_num, p.fps_den));
    b.ue(0);         // seq_parameter_set_id
    if (p.transform8x8) {   // profile_idc 100 carries the chroma format / bit depth fields
        b.ue(1);            // chroma_format_idc: 4:2:0
        b.ue(0); b.ue(0);   // bit_depth_luma_minus8, bit_depth_chroma_minus8
        b.put(1, 0);        // qpprime_y_zero_transform_bypass_flag
        b.put(1, 0);        // seq_scaling_matrix_present_flag
    }
    b.ue(4);         // log2_max_frame_num_minus4
    b.ue(2);         // pic_order_cnt_type
    b.ue(1);         // max_num_ref_frames
    b.put(1, 0);     // gaps_in_frame_num_value_allowed_flag
    b.ue((uint32_t)(mbw - 1));
    b.ue((uint32_t)(mbh - 1));
    b.put(1, 1);     // frame_mbs_only_flag
    b.put(1, 1);     // direct_8x8_inference_flag
    const int cr = 16 * mbw - p.width, cb = 16 * mbh - p.height;
    if (cr || cb) { b.put(1, 1); b.ue(0); b.ue((uint32_t)(cr / 2)); b.ue(0); b.ue((uint32_t)(cb / 2)); }
    else b.put(1, 0);
    b.put(1, 1);     // vui_parameters_present_flag
    b.put(1, 0);     // aspect_ratio_info_present_flag
    b.put(1, 0);     // overscan_info_present_flag
    b.put(1, 0);     // video_signal_type_present_flag
    b.put(1, 0);     // chroma_loc_info_present_flag
    b.put(1, 1);     // timing_info_present_flag
    b.put32((uint32_t)p.fps_den);
    b.put32((uint32_t)p.fps_num * 2);
    b.put(1, 1);     // fixed_frame_rate_flag
    b.put(1, 0);     // nal_hrd_parameters_present_flag
    b.put(1, 0);     // vcl_hrd_parameters_present_flag
    b.put(1, 0);     // pic_struct_present_flag
    b.put(1, 1);     // bitstream_restriction_flag
    b.put(1, 1);     // motion_vectors_over_pic_boundaries_flag
    b.ue(0);         // max_bytes_per_pic_denom
    b.ue(0);         // max_bits_per_mb_denom
    b.ue(9);         // log2_max_mv_length_horizontal
    b.ue(9);         // log2_max_mv_length_vertical
    b.ue(0);         // max_num_reorder_frames
    b.ue(1);         // max_dec_frame_buffering
    b.trailing();
    return nal_escape(3, 7, b.bytes());
}

std::vector<uint8_t> make_pps_nal(const vcpenc_params& p) {
    BitWriter b;
    b.ue(0); b.ue(0);
    b.put(1, p.entropy ? 1 : 0);   // entropy_coding_mode_flag: 0 CAVLC, 1 CABAC
    b.put(1, 0);     // bottom_field_pic_order_in_frame_present_flag
    b.ue(0);         // num_slice_groups_minus1
    b.ue(0); b.ue(0);
    b.put(1, 0);     // weighted_pred_flag
    b.put(2, 0);     // weighted_bipred_idc
    b.se(0);         // pic_init_qp_minus26
    b.se(0);         // pic_init_qs_minus26
    b.se(0);         // chroma_qp_index_offset
    b.put(1, 1);     // deblocking_filter_control_present_flag
    b.put(1, 0);     // constrained_intra_pred_flag
    b.put(1, 0);     // redundant_pic_cnt_present_flag
    if (p.transform8x8) {
        b.put(1, 1); // transform_8x8_mode_flag
        b.put(1, 0); // pic_scaling_matrix_present_flag
        b.se(0);     // second_chroma_qp_index_offset
    }
    b.trailing();
    return nal_escape(3, 8, b.bytes());
}

}  // namespace vcp
