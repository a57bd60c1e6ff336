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

// ---- HEVC (H.265 7.3.2.1 - 7.3.2.3, 7.3.3, E.2.1) ------------------------------------------------------
// Main profile, 16x16 coding tree blocks that are never split, 8x8 / 4x4 transform blocks, one reference
// picture (short-term RPS {-1}), no temporal vector prediction, deblocking unless deblock_idc = 1, no SAO.

int hevc_level_idc_for(int cw, int ch, int fps_num, int fps_den) {
    // table A.8: general_level_idc = 30 x level; MaxLumaPs, MaxLumaSr
    static const struct { int idc; long ps; double sr; } L[] = {
        {30, 36864, 552960},        {60, 122880, 3686400},       {63, 245760, 7372800},      {90, 552960, 16588800},
        {93, 983040, 33177600},     {120, 2228224, 66846720},    {123, 2228224, 133693440},  {150, 8912896, 267386880},
        {153, 8912896, 534773760},  {156, 8912896, 1069547520},  {180, 35651584, 1069547520},
        {183, 35651584, 2139095040.0}, {186, 35651584, 4278190080.0}};
    const long ps = (long)cw * ch;
    const double sr = (double)ps * fps_num / (fps_den > 0 ? fps_den : 1);
    for (const auto& l : L)
        if (ps <= l.ps && sr <= l.sr) return l.idc;
    return 186;
}

static void hevc_profile_tier_level(BitWriter& b, const vcpenc_params& p) {
    const int cw = (p.width + 15) / 16 * 16, ch = (p.height + 15) / 16 * 16;
    b.put(2, 0); b.put(1, 0); b.put(5, 1);      // general_profile_space, tier, profile_idc: Main
    b.put32(0x60000000u);                       // compatibility flags: Main, Main 10
    b.put(1, 1); b.put(1, 0); b.put(1, 0); b.put(1, 1);   // progressive, interlaced, non-packed, frame-only
    b.put(22, 0); b.put(22, 0);                 // reserved zero bits (43) and general_inbld_flag
    b.put(8, (uint32_t)hevc_level_idc_for(cw, ch, p.fps_num, p.fps_den));
}

std::vector<uint8_t> make_hevc_vps_nal(const vcpenc_params& p) {
    BitWriter b;
    b.put(4, 0);            // vps_video_parameter_set_id
    b.put(1, 1); b.put(1, 1);   // base layer internal / available
    b.put(6, 0);            // vps_max_layers_minus1
    b.put(3, 0);            // vps_max_sub_layers_minus1
    b.put(1, 1);            // vps_temporal_id_nesting_flag
    b.put(16, 0xffff);
    hevc_profile_tier_level(b, p);
    b.put(1, 1);            // vps_sub_layer_ordering_info_present_flag
    b.ue(1); b.ue(0); b.ue(0);  // max_dec_pic_buffering_minus1, max_num_reorder_pics, max_latency_increase_plus1
    b.put(6, 0);            // vps_max_layer_id
    b.ue(0);                // vps_num_layer_sets_minus1
    b.put(1, 0);            // vps_timing_info_present_flag
    b.put(1, 0);            // vps_extension_flag
    b.trailing();
    return hevc_nal_escape(32, b.bytes());
}

std::vector<uint8_t> make_hevc_sps_nal(const vcpenc_params& p) {
    const int cw = (p.width + 15) / 16 * 16, ch = (p.height + 15) / 16 * 16;
    BitWriter b;
    b.put(4, 0); b.put(3, 0); b.put(1, 1);      // vps id, max_sub_layers_minus1, temporal_id_nesting
    hevc_profile_tier_level(b, p);
    b.ue(0);                // sps_seq_parameter_set_id
    b.ue(1);                // chroma_format_idc
    b.ue((uint32_t)cw); b.ue((uint32_t)ch);
    const int cr = cw - p.width, cb = ch - p.height;
    if (cr || cb) { b.put(1, 1); b.ue(0); b.ue((uint32_t)(cr / 2)); b.ue(0); b.ue((uint32_t)(cb / 2)); }
    else b.put(1, 0);
    b.ue(0); b.ue(0);       // bit_depth_luma_minus8, bit_depth_chroma_minus8
    b.ue(4);                // log2_max_pic_order_cnt_lsb_minus4
    b.put(1, 1); b.ue(1); b.ue(0); b.ue(0);     // sub-layer ordering info
    b.ue(1);                // log2_min_luma_coding_block_size_minus3 (16)
    b.ue(0);                // log2_diff_max_min_luma_coding_block_size
    b.ue(0);                // log2_min_luma_transform_block_size_minus2 (4)
    b.ue(1);                // log2_diff_max_min_luma_transform_block_size (8)
    b.ue(0); b.ue(0);       // max_transform_hierarchy_depth_inter / intra
    b.put(1, 0);            // scaling_list_enabled_flag
    b.put(1, 0);            // amp_enabled_flag
    b.put(1, p.hevc_sao ? 1 : 0);   // sample_adaptive_offset_enabled_flag
    b.put(1, 0);            // pcm_enabled_flag
    b.ue(1);                // num_short_term_ref_pic_sets
    b.ue(1); b.ue(0); b.ue(0); b.put(1, 1);     // one negative picture: delta_poc_s0_minus1 0, used
    b.put(1, 0);            // long_term_ref_pics_present_flag
    b.put(1, 0);            // sps_temporal_mvp_enabled_flag
    b.put(1, 0);            // strong_intra_smoothing_enabled_flag
    b.put(1, 1);            // vui_parameters_present_flag
    b.put(1, 0); b.put(1, 0); b.put(1, 0); b.put(1, 0);   // aspect ratio, overscan, video signal, chroma loc
    b.put(1, 0); b.put(1, 0); b.put(1, 0);                // neutral chroma, field_seq, frame_field_info
    b.put(1, 0);            // default_display_window_flag
    b.put(1, 1);            // vui_timing_info_present_flag
    b.put32((uint32_t)p.fps_den); b.put32((uint32_t)p.fps_num);
    b.put(1, 0);            // vui_poc_proportional_to_timing_flag
    b.put(1, 0);            // vui_hrd_parameters_present_flag
    b.put(1, 0);            // bitstream_restriction_flag
    b.put(1, 0);            // sps_extension_present_flag
    b.trailing();
    return hevc_nal_escape(33, b.bytes());
}

std::vector<uint8_t> make_hevc_pps_nal(const vcpenc_params& p) {
    BitWriter b;
    b.ue(0); b.ue(0);       // pps id, sps id
    b.put(1, 0);            // dependent_slice_segments_enabled_flag
    b.put(1, 0);            // output_flag_present_flag
    b.put(3, 0);            // num_extra_slice_header_bits
    b.put(1, 0);            // sign_data_hiding_enabled_flag
    b.put(1, 0);            // cabac_init_present_flag
    b.ue(0); b.ue(0);       // num_ref_idx_l0/l1_default_active_minus1
    b.se(0);                // init_qp_minus26
    b.put(1, 0);            // constrained_intra_pred_flag
    b.put(1, 0);            // transform_skip_enabled_flag
    b.put(1, 0);            // cu_qp_delta_enabled_flag
    b.se(0); b.se(0);       // pps_cb_qp_offset, pps_cr_qp_offset
    b.put(1, 0);            // pps_slice_chroma_qp_offsets_present_flag
    b.put(1, 0); b.put(1, 0);   // weighted_pred_flag, weighted_bipred_flag
    b.put(1, 0);            // transquant_bypass_enabled_flag
    b.put(1, 0);            // tiles_enabled_flag
    b.put(1, 0);            // entropy_coding_sync_enabled_flag
    b.put(1, 0);            // pps_loop_filter_across_slices_enabled_flag
    if (p.deblock_idc == 1) {
        b.put(1, 1);        // deblocking_filter_control_present_flag
        b.put(1, 0);        // deblocking_filter_override_enabled_flag
        b.put(1, 1);        // pps_deblocking_filter_disabled_flag
    } else b.put(1, 0);     // no control syntax: deblocking on, beta / tc offsets 0
    b.put(1, 0);            // pps_scaling_list_data_present_flag
    b.put(1, 0);            // lists_modification_present_flag
    b.ue(0);                // log2_parallel_merge_level_minus2
    b.put(1, 0);            // slice_segment_header_extension_present_flag
    b.put(1, 0);            // pps_extension_present_flag
    b.trailing();
    return hevc_nal_escape(34, b.bytes());
}

}  // namespace vcp
