"""ctypes binding of include/selenite_b200.h. Loading fails loudly: there is no Python or CPU fallback."""
import ctypes as C
import os

from . import build as _build

MAX_STAGES = 4
_lib = None


class Config(C.Structure):
    _fields_ = [("channels", C.c_uint32), ("fs", C.c_uint32), ("device", C.c_int32), ("chain", C.c_uint32)]


class RxF32Params(C.Structure):
    _fields_ = [("fft_len", C.c_uint32), ("hop", C.c_uint32), ("agc_block", C.c_uint32), ("n_stages", C.c_uint32),
                ("biquad", C.c_float * (5 * MAX_STAGES)),
                ("agc_target", C.c_float), ("agc_decay", C.c_float), ("agc_floor", C.c_float), ("agc_gmax", C.c_float)]


# every symbol include/selenite_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "slb_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "slb_destroy": (None, [_P]),
    "slb_last_error": (C.c_char_p, [_P]),
    "slb_version": (C.c_char_p, []),
    "slb_default_rx_f32_params": (C.c_int, [C.c_uint32, C.POINTER(RxF32Params)]),
    "slb_default_mask": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint8, _P]),
    "slb_set_rx_f32_params": (C.c_int, [_P, C.POINTER(RxF32Params)]),
    "slb_get_rx_f32_params": (C.c_int, [_P, C.POINTER(RxF32Params)]),
    "slb_set_mask": (C.c_int, [_P, C.c_uint8, _P]),
    "slb_get_mask": (C.c_int, [_P, C.c_uint8, _P]),
    "SLB_DSP_Init": (C.c_int, [_P]),
    "SLB_DSP_Set_RX": (C.c_int, [_P]),
    "SLB_DSP_Set_TX": (C.c_int, [_P]),
    "SLB_DSP_Set_Mode": (C.c_int, [_P, C.c_uint8]),
    "SLB_DSP_Set_Mode_Channel": (C.c_int, [_P, C.c_uint32, C.c_uint8]),
    "SLB_DSP_In_Buff_Write": (C.c_int, [_P, _P, C.c_uint16]),
    "SLB_DSP_In_Buff_Read": (C.c_int, [_P, _P, C.c_uint32]),
    "SLB_DSP_Out_Buff_Write": (C.c_int, [_P, _P, C.c_uint32]),
    "SLB_DSP_Out_Buff_Read": (C.c_int, [_P, _P, C.c_uint16]),
    "SLB_DSP_Out_Buff_Mute": (C.c_int, [_P]),
    "slb_ring_get_ptrs": (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint32 * 3)]),
    "slb_ring_get_iq": (C.c_int, [_P, C.c_int, _P, _P]),
    "slb_rx_process_device": (C.c_int, [_P, _P, _P, C.c_uint32, _P]),
    "slb_rx_process_host": (C.c_int, [_P, _P, _P, C.c_uint32]),
    "slb_rx_set_debug_taps": (C.c_int, [_P, _P, _P]),
    "slb_state_size": (C.c_int, [_P, C.POINTER(C.c_size_t)]),
    "slb_state_save": (C.c_int, [_P, _P, C.c_size_t]),
    "slb_state_load": (C.c_int, [_P, _P, C.c_size_t]),
    "slb_ring_plan_write": (C.c_uint32, [C.c_uint32, C.c_int, C.POINTER(C.c_uint32 * 3), C.c_uint32]),
    "slb_ring_plan_read": (C.c_uint32, [C.c_uint32, C.c_int, C.POINTER(C.c_uint32 * 3), C.c_uint32]),
    "slb_biquad_scan_tables": (C.c_int, [_P, _P, _P]),
    "slb_kernel_launches": (C.c_uint64, [_P]),
    "slb_sync": (C.c_int, [_P]),
    "DSP_Init": (None, []),
    "DSP_Set_RX": (None, []),
    "DSP_Set_TX": (None, []),
    "DSP_Set_Mode": (None, [C.c_uint8]),
    "DSP_In_Buff_Write": (None, [_P, C.c_uint16]),
    "DSP_In_Buff_Read": (None, [_P, C.c_uint32]),
    "DSP_Out_Buff_Write": (None, [_P, C.c_uint32]),
    "DSP_Out_Buff_Read": (None, [_P, C.c_uint16]),
    "DSP_Out_Buff_Mute": (None, []),
    "slb_dropin_status": (C.c_int, []),
    "slb_dropin_ctx": (_P, []),
}


def lib_path():
    return _build.LIB_PATH


def load():
    """Load lib/libselenite_b200.so (building it first when nvcc is present and the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if _build.is_stale():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this box: use the prebuilt library if there is one
            if not os.path.exists(path):
                raise RuntimeError("selenite_lite_b200: CUDA library missing and cannot be built (%s); "
                                   "there is no CPU fallback" % exc)
    if not os.path.exists(path):
        raise RuntimeError("selenite_lite_b200: %s not found; run __graft_entry__.build(). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
