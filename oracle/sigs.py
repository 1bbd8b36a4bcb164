"""ctypes signatures for liboracle.so (test infrastructure)."""
import ctypes as C

v = C.c_void_p
i = C.c_int

SIGS = {
    "oracle_i420_to_rgb32": (None, [v, v, i, i]),
    "oracle_half_rgb": (None, [v, v, i, i]),
    "oracle_flip_rgb": (None, [v, v, i, i, i, i]),
    "oracle_convert_to_i420": (i, [v, C.c_size_t, v, i, v, i, v, i, i, i, C.c_uint32]),
    "oracle_mjpg_to_i420": (i, [v, C.c_size_t, v, i, v, i, v, i, i, i]),
    "oracle_mjpg_planes": (i, [v, C.c_size_t, v, v, v, v, v]),
}


def bind(lib) -> None:
    for name, (res, args) in SIGS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            continue
        fn.restype = res
        fn.argtypes = args


class OrcCu(C.Structure):
    _fields_ = [("mvx", C.c_int16), ("mvy", C.c_int16), ("log2_size", C.c_uint8), ("pred_mode", C.c_uint8),
                ("intra_mode", C.c_uint8), ("cbf", C.c_uint8), ("skip", C.c_uint8), ("merge_idx", C.c_uint8),
                ("mvp_idx", C.c_uint8), ("qp", C.c_uint8), ("ref_idx", C.c_uint8), ("chroma_mode", C.c_uint8),
                ("tu_log2", C.c_uint8), ("flags", C.c_uint8)]


class OrcEncCfg(C.Structure):
    _fields_ = [("width", i), ("height", i), ("qp", i), ("intra_period", i), ("search_range", i),
                ("deblock", i), ("hash_sei", i), ("qp_delta", i),
                ("mv_edges", i), ("more_tiles", i), ("raw_slice_data", i), ("no_wpp", i), ("subme_satd", i), ("sao", i), ("tile_cols", i),
                ("tr_depth", i), ("cabac_init", i), ("refs", i), ("tmvp", i), ("me_coarse", i), ("intra_in_p", i), ("fps_num", i), ("fps_den", i),
                ("intra_satd", i), ("tu4", i), ("intra_sizes", i), ("chroma_modes", i), ("sign_hiding", i), ("strong_intra", i),
                ("cb_qp_offset", i), ("cr_qp_offset", i), ("beta_offset_div2", i), ("tc_offset_div2", i), ("tile_rows", i), ("vaq", i), ("scaling_list", i), ("conf_right", i), ("conf_bottom", i)]


SIGS.update({
    "orc_enc_open": (v, [C.POINTER(OrcEncCfg)]),
    "orc_enc_close": (None, [v]),
    "orc_set_threads": (None, [i]),
    "orc_max_threads": (i, []),
    "orc_enc_encode": (i, [v, v, v, i]),
    "orc_enc_set_ctu_dqp": (i, [v, v]),
    "orc_enc_set_qp": (i, [v, i]),
    "orc_scaling_table": (None, [i, v]),
    "orc_vaq_offsets": (None, [v, i, i, i, v]),
    "orc_tiled_open": (v, [C.POINTER(OrcEncCfg), i]),
    "orc_tiled_open2": (v, [C.POINTER(OrcEncCfg), i, i]),
    "orc_tiled_close": (None, [v]),
    "orc_tiled_encode": (i, [v, v, v, i]),
    "orc_tiled_recon": (v, [v]),
    "orc_enc_recon": (v, [v]),
    "orc_enc_recon_predeblock": (v, [v]),
    "orc_enc_cu_map": (v, [v]),
    "orc_enc_levels": (v, [v]),
    "orc_enc_last_was_idr": (i, [v]),
    "orc_enc_bins": (C.c_ulonglong, [v]),
    "orc_sad": (C.c_uint32, [v, i, v, i, i, i]),
    "orc_satd": (C.c_uint32, [v, i, v, i, i, i]),
    "orc_fdct": (None, [v, v, i]),
    "orc_idct": (None, [v, v, i]),
    "orc_fdst4": (None, [v, v]),
    "orc_idst4": (None, [v, v]),
    "orc_quant": (i, [v, v, i, i, i]),
    "orc_dequant": (None, [v, v, i, i]),
    "orc_chroma_qp": (i, [i]),
    "orc_intra_predict": (None, [v, i, i, i, v, i]),
    "orc_mc_luma": (None, [v, i, i, i, i, i, i, i, i, i, v, i]),
    "orc_mc_chroma": (None, [v, i, i, i, i, i, i, i, i, i, v, i]),
    "orc_deblock_luma_segment": (None, [v, i, i, i, i]),
    "orc_deblock_chroma_segment": (None, [v, i, i, i, i]),
    "orc_dct_coef": (i, [i, i, i]),
})
