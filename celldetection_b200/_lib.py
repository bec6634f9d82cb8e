"""ctypes binding of ``libcpn_b200.so`` (the C ABI declared in ``include/cpn_b200.h``).

This is the only place the package touches native code.  There is no CPU fallback: if the shared object is missing
or a call fails, a ``RuntimeError`` is raised.  Tensors cross the boundary as raw ``data_ptr()`` device pointers plus
explicit sizes and the current CUDA stream handle.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcpn_b200.so')

ABI_VERSION = 5

# ---- enums (mirror include/cpn_b200.h) -------------------------------------------------------------------------------
DT_F32, DT_F16, DT_U8, DT_F16X2, DT_F16F8, DT_U16 = 0, 1, 2, 3, 4, 5
OP_PREP, OP_CONV, OP_MAXPOOL, OP_UPSAMPLE, OP_BILINEAR, OP_PROJ = range(6)
IN_F32_NCHW, IN_U8_NCHW, IN_U8_NHWC = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SCALED_TANH, ACT_SIGMOID = 0, 1, 2, 3
ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1
CONV_UP2 = 1


class View(ctypes.Structure):
    _fields_ = [('offset', ctypes.c_int64), ('n', ctypes.c_int32), ('h', ctypes.c_int32), ('w', ctypes.c_int32),
                ('c', ctypes.c_int32), ('pitch', ctypes.c_int32), ('dtype', ctypes.c_int32),
                ('lo_delta', ctypes.c_int32), ('fp8_exp', ctypes.c_int32)]


class Op(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('engine', ctypes.c_int32), ('src', View), ('dst', View), ('res', View),
                ('w_offset', ctypes.c_int64), ('b_offset', ctypes.c_int64),
                ('r', ctypes.c_int32), ('s', ctypes.c_int32), ('stride', ctypes.c_int32), ('pad', ctypes.c_int32),
                ('kslab', ctypes.c_int32), ('slab_mode', ctypes.c_int32), ('act', ctypes.c_int32),
                ('act_scale', ctypes.c_float), ('proj_cin_off', ctypes.c_int32), ('proj_cin', ctypes.c_int32),
                ('out_binding', ctypes.c_int32), ('fuse_next', ctypes.c_int32), ('acc_scale', ctypes.c_float),
                ('flags', ctypes.c_int32)]


class SelectParams(ctypes.Structure):
    """cpn_select_params_t"""
    _fields_ = [('logits', ctypes.c_void_p), ('lower', ctypes.c_void_p), ('upper', ctypes.c_void_p),
                ('uncertainty', ctypes.c_void_p), ('channels', ctypes.c_int32), ('use_certainty', ctypes.c_int32),
                ('thresh', ctypes.c_float), ('certainty_limit', ctypes.c_float)]


# name -> (restype, argtypes); every symbol include/cpn_b200.h declares
_P, _I, _I64, _F, _SZ = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
SYMBOLS = {
    'cpn_abi_version': (_I, []),
    'cpn_last_error': (ctypes.c_char_p, []),
    'cpn_launch_count': (_I64, []),
    'cpn_device_info': (_I, [ctypes.c_char_p, _I, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I)]),
    'cpn_plan_create': (_I, [ctypes.POINTER(Op), _I, _P, _SZ, _P, _SZ, _P, ctypes.POINTER(_P)]),
    'cpn_plan_forward': (_I, [_P, _P, _I, ctypes.POINTER(_P), _I, _P]),
    'cpn_plan_forward_range': (_I, [_P, _I, _I, _P, _I, ctypes.POINTER(_P), _I, _P]),
    'cpn_plan_num_launches': (_I, [_P]),
    'cpn_plan_set_active_rows': (_I, [_P, _I64]),
    'cpn_plan_run_op': (_I, [_P, _I, _P, _I, ctypes.POINTER(_P), _I, _P]),
    'cpn_plan_destroy': (None, [_P]),
    'cpn_conv2d': (_I, [ctypes.POINTER(Op), _P, _P, _P, _P, _P]),
    'cpn_select_workspace_bytes': (_SZ, [_I64]),
    'cpn_select_count': (_I, [_P, _P, _P, _I64, _F, _P, _P, _P]),
    'cpn_select_write': (_I, [_P, _P, _P, _I, _I64, _F, _P, _P, _P, _I64, _P, _P]),
    'cpn_select_count_ex': (_I, [ctypes.POINTER(SelectParams), _I64, _P, _P, _P]),
    'cpn_select_write_ex': (_I, [ctypes.POINTER(SelectParams), _I, _I64, _P, _P, _P, _P, _I64, _P, _P]),
    'cpn_resize_bilinear': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _P]),
    'cpn_nms_weights': (_I, [_P, _P, _P, _I64, _P, _P]),
    'cpn_decode_refine_buckets': (_I, [_P, _I64, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P, _P, _P, _P, _P,
                                       _P, _P, _P, _P]),
    'cpn_decode_refine': (_I, [_P, _I64, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P]),
    'cpn_decode_refine_rows': (_I, [_P, _I64, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P, _P, _P, _P, _P,
                                    _P, _P, _P, _P]),
    'cpn_gather_patches': (_I, [_P, ctypes.POINTER(View), _P, _I64, _I, _P, ctypes.POINTER(View), _P]),
    'cpn_fouriers2contours': (_I, [_P, _P, _I64, _I, _I, _P, _P, _P, _P]),
    'cpn_nms_workspace_bytes': (_SZ, [_I64, _I]),
    'cpn_nms_segments': (_I, [_P, _P, _P, _I, _I64, _F, _I, _P, _P, _P, _P]),
    'cpn_nms_grid_workspace_bytes': (_SZ, [_I64]),
    'cpn_nms_grid': (_I, [_P, _P, _I64, _F, _P, _P, _P, ctypes.POINTER(_I), _P]),
    'cpn_box_votes_workspace_bytes': (_SZ, [_I64]),
    'cpn_box_votes': (_I, [_P, _I64, _F, _P, _P, _P]),
    'cpn_border_filter': (_I, [_P, _P, _P, _I64, _I, _F, _P, _P]),
    'cpn_contours2labels_workspace_bytes': (_SZ, [_I64, _I]),
    'cpn_contours2labels': (_I, [_P, _I64, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P]),
    'cpn_resolve_label_channels_workspace_bytes': (_SZ, [_I, _I]),
    'cpn_resolve_label_channels': (_I, [_P, _I, _I, _I, _I, _P, _P, ctypes.POINTER(_I), _P]),
    'cpn_gather_rows': (_I, [_P, _I64, _P, _I64, _P, _P]),
    'cpn_unshuffle2': (_I, [_P, _I, _I, _I, _I, _P, _P]),
    'cpn_copy_window': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'cpn_label_props': (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    'cpn_histogram': (_I, [_P, _I, _I64, _P, _P]),
    'cpn_apply_lut': (_I, [_P, _I, _I64, _P, _P, _P]),
    'cpn_rgb2gray': (_I, [_P, _I, _I64, _I, _P, _P, _P]),
}

_lib = None


def load():
    """Load the shared object (once).  Raises if it is missing -- there is deliberately no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} not found: build it with `python -m celldetection_b200.build` (nvcc, sm_100a). '
            'celldetection_b200 has no CPU or PyTorch fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.cpn_abi_version() != ABI_VERSION:
        raise RuntimeError(f'libcpn_b200.so ABI {lib.cpn_abi_version()} != expected {ABI_VERSION}; rebuild')
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().cpn_last_error().decode(errors='replace')
        raise RuntimeError(f'cpn_b200 {what} failed: {msg}')


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def addr(t):
    """Device address of a tensor as a plain int (None -> NULL), for pointer fields of ctypes structures."""
    return None if t is None else int(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class nvtx_range:
    """NVTX range around a stage of the path when ``CPN_NVTX=1`` (nsys / ncu --nvtx timelines: ``cpn.plan``, ``cpn.post``,
    ``cpn.tiles``, ``cpn.exchange``, ``cpn.stitch``, ``cpn.preprocess``); a no-op otherwise."""
    enabled = os.environ.get('CPN_NVTX', '0') == '1'

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if nvtx_range.enabled:
            import torch
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if nvtx_range.enabled:
            import torch
            torch.cuda.nvtx.range_pop()
        return False


def launch_count():
    return int(load().cpn_launch_count())


def device_info():
    lib = load()
    name = ctypes.create_string_buffer(256)
    sm, ma, mi = _I(), _I(), _I()
    check(lib.cpn_device_info(name, 256, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi)), 'device_info')
    return dict(name=name.value.decode(), sm_count=sm.value, cc=(ma.value, mi.value))
