"""ctypes binding of the C ABI declared in include/lh2b.h. No fallback: a missing library raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libRenderCore_B200.so")


class LibraryMissing(RuntimeError):
    pass


_lib = None
_c = ctypes
_vp, _ip, _fp = _c.c_void_p, _c.c_int, _c.POINTER(_c.c_float)

SIGNATURES = {
    "lh2b_create": ([_c.POINTER(_vp), _ip], _ip),
    "lh2b_destroy": ([_vp], _ip),
    "lh2b_last_error": ([], _c.c_char_p),
    "lh2b_set_target": ([_vp, _ip, _ip, _ip], _ip),
    "lh2b_setting": ([_vp, _c.c_char_p, _c.c_float], _ip),
    "lh2b_set_probe_pos": ([_vp, _ip, _ip], _ip),
    "lh2b_set_textures": ([_vp, _vp, _ip], _ip),
    "lh2b_set_materials": ([_vp, _vp, _ip], _ip),
    "lh2b_set_lights": ([_vp, _vp, _ip, _vp, _ip, _vp, _ip, _vp, _ip], _ip),
    "lh2b_set_sky": ([_vp, _vp, _ip, _ip, _vp], _ip),
    "lh2b_set_geometry": ([_vp, _ip, _vp, _ip, _ip, _vp], _ip),
    "lh2b_set_geometry_device": ([_vp, _ip, _vp, _ip, _ip, _vp], _ip),
    "lh2b_set_skin": ([_vp, _ip, _vp, _vp, _ip], _ip),
    "lh2b_set_pose": ([_vp, _ip, _vp, _ip], _ip),
    "lh2b_set_morph_targets": ([_vp, _ip, _vp, _vp, _ip, _ip], _ip),
    "lh2b_set_morph_weights": ([_vp, _ip, _vp, _ip], _ip),
    "lh2b_read_geometry": ([_vp, _ip, _vp, _vp], _ip),
    "lh2b_set_instance": ([_vp, _ip, _ip, _vp], _ip),
    "lh2b_finalize_instances": ([_vp], _ip),
    "lh2b_render": ([_vp, _vp, _ip, _ip], _ip),
    "lh2b_wait_for_render": ([_vp], _ip),
    "lh2b_get_stats": ([_vp, _vp], _ip),
    "lh2b_read_pixels": ([_vp, _vp], _ip),
    "lh2b_read_pixels_async": ([_vp, _vp], _ip),
    "lh2b_wait_read_pixels": ([_vp], _ip),
    "lh2b_read_accumulator": ([_vp, _vp], _ip),
    "lh2b_accumulator_device_ptr": ([_vp, _c.POINTER(_vp), _c.POINTER(_ip)], _ip),
    "lh2b_finalize_external": ([_vp, _vp, _ip], _ip),
    "lh2b_snapshot_accumulator": ([_vp, _vp], _ip),
    "lh2b_host_bvh_build": ([_vp, _ip, _vp, _ip, _vp, _ip, _c.POINTER(_ip)], _ip),
    "lh2b_set_row_band": ([_vp, _ip, _ip], _ip),
    "lh2b_set_row_band_strided": ([_vp, _ip, _ip, _ip], _ip),
    "lh2b_tile_handle_bytes": ([], _ip),
    "lh2b_tile_layout": ([_ip, _ip, _c.c_float, _ip, _c.POINTER(_ip), _c.POINTER(_ip), _c.POINTER(_ip)], _ip),
    "lh2b_tile_shard_layout": ([_ip, _ip, _ip, _ip, _c.POINTER(_ip), _c.POINTER(_ip), _c.POINTER(_ip), _c.POINTER(_ip)], _ip),
    "lh2b_tile_rows_inside": ([_ip, _ip, _ip, _ip, _ip, _c.POINTER(_ip), _c.POINTER(_ip)], _ip),
    "lh2b_tile_create": ([_vp, _ip, _ip, _c.POINTER(_vp)], _ip),
    "lh2b_tile_export": ([_vp, _vp], _ip),
    "lh2b_tile_import": ([_vp, _vp], _ip),
    "lh2b_tile_frame": ([_vp], _ip),
    "lh2b_tile_wait": ([_vp], _ip),
    "lh2b_tile_rows": ([_vp, _c.POINTER(_ip), _c.POINTER(_ip), _c.POINTER(_ip)], _ip),
    "lh2b_tile_destroy": ([_vp], _ip),
    "lh2b_gather_handle_bytes": ([], _ip),
    "lh2b_gather_create": ([_vp, _ip, _ip, _c.POINTER(_vp)], _ip),
    "lh2b_gather_export": ([_vp, _vp], _ip),
    "lh2b_gather_import": ([_vp, _vp], _ip),
    "lh2b_gather_frame": ([_vp, _ip, _vp], _ip),
    "lh2b_gather_wait": ([_vp], _ip),
    "lh2b_gather_join": ([_vp, _vp], _ip),
    "lh2b_gather_image_device_ptr": ([_vp, _c.POINTER(_vp)], _ip),
    "lh2b_gather_destroy": ([_vp], _ip),
    "lh2b_finalize_external_on": ([_vp, _vp, _ip, _vp, _vp], _ip),
    "lh2b_set_sample_shard": ([_vp, _ip, _ip], _ip),
    "lh2b_trace_rays": ([_vp, _vp, _vp, _ip, _vp], _ip),
    "lh2b_trace_shadow_rays": ([_vp, _vp, _vp, _ip, _vp], _ip),
    "lh2b_trace_rays_device": ([_vp, _vp, _vp, _ip, _vp, _ip, _fp], _ip),
    "lh2b_trace_shadow_rays_device": ([_vp, _vp, _vp, _ip, _vp, _ip, _fp], _ip),
    "lh2b_trace_stats_enable": ([_vp, _ip], _ip),
    "lh2b_trace_stats_read": ([_vp, _vp, _ip], _ip),
    "lh2b_stream": ([_vp, _c.POINTER(_vp)], _ip),
    "lh2b_get_frame_stats": ([_vp, _vp], _ip),
    "lh2b_get_bvh_stats": ([_vp, _ip, _vp], _ip),
    "lh2b_handle_of": ([_vp], _vp),
    "lh2b_present_gl": ([_vp, _c.c_uint], _ip),
    "lh2b_filter_chain": ([_vp, _vp], _ip),
    "lh2b_shade_paths_time": ([_vp, _ip, _ip, _vp, _vp, _vp, _vp, _c.c_uint, _c.c_uint, _ip, _ip, _vp], _ip),
    "lh2b_read_filter_buffers": ([_vp, _vp, _vp, _vp, _vp], _ip),
    "lh2b_read_filter_history": ([_vp, _vp, _vp, _vp, _vp, _vp], _ip),
    "lh2b_debug_read_table": ([_vp, _c.c_char_p, _vp, _c.c_size_t, _c.POINTER(_c.c_size_t)], _ip),
    "lh2b_shade_paths": ([_vp, _ip, _ip, _vp, _vp, _vp, _vp, _c.c_uint, _c.c_uint, _ip, _vp, _vp, _vp, _c.POINTER(_ip),
                          _vp, _vp, _vp, _c.POINTER(_ip), _vp], _ip),
    "CreateCore": ([], _vp),
}


def load_library(path=None):
    """Load libRenderCore_B200.so and attach argument types. Raises LibraryMissing if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise LibraryMissing(f"{p} not found: build it with `make -C lighthouse2_b200/csrc` "
                             "(or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes, fn.restype = argtypes, restype
    if path is None:
        _lib = lib
    return lib
