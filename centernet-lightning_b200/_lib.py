"""ctypes binding of libcnl_b200.so (the C ABI declared in include/cnl_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# CNL_LIB points at another build of the SAME library (A/B timing of kernel variants); it is never a fallback
LIB_PATH = os.environ.get("CNL_LIB") or os.path.join(_HERE, "libcnl_b200.so")

EXPORTS = (
    "cnl_last_error", "cnl_version", "cnl_compiled_sm",
    "cnl_decode_workspace_bytes", "cnl_decode_workspace_bytes_k", "cnl_decode_detections", "cnl_decode_detections_packed", "cnl_gather_boxes", "cnl_sigmoid", "cnl_boxes_xyxy_to_xywh",
    "cnl_normalize_images_u8", "cnl_track_workspace_bytes", "cnl_track_cost_matrices",
    "cnl_engine_create", "cnl_engine_destroy", "cnl_engine_arena_bytes", "cnl_engine_buffer_offset", "cnl_engine_op_form",
    "cnl_engine_upload", "cnl_engine_forward", "cnl_engine_forward_act", "cnl_engine_read_buffer", "cnl_engine_write_buffer",
)


class ConvDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("src", C.c_int), ("dst", C.c_int), ("cin", C.c_int), ("cout", C.c_int),
                ("ksize", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("relu", C.c_int),
                ("src_c_off", C.c_int), ("dst_c_off", C.c_int), ("residual", C.c_int), ("residual_up", C.c_int),
                ("kh", C.c_int), ("kw", C.c_int), ("pad_h", C.c_int), ("pad_w", C.c_int),
                ("dst_up", C.c_int), ("dst_phase", C.c_int),
                ("weight_host", C.c_void_p), ("bias_host", C.c_void_p),
                ("src2", C.c_int), ("src3", C.c_int), ("scale0", C.c_float), ("scale1", C.c_float), ("scale2", C.c_float),
                ("resize", C.c_int)]


class BufferDesc(C.Structure):
    _fields_ = [("channels", C.c_int), ("stride", C.c_int), ("fp32_nchw", C.c_int)]


class CnlError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CnlError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU or PyTorch fallback for the cnl_b200 hot path)")
    lib = C.CDLL(LIB_PATH)
    lib.cnl_last_error.restype = C.c_char_p
    lib.cnl_version.restype = C.c_int
    lib.cnl_compiled_sm.restype = C.c_int
    lib.cnl_decode_workspace_bytes.restype = C.c_size_t
    lib.cnl_decode_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.cnl_decode_detections.restype = C.c_int
    lib.cnl_decode_detections.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_size_t, C.c_void_p]
    lib.cnl_decode_workspace_bytes_k.restype = C.c_size_t
    lib.cnl_decode_workspace_bytes_k.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    lib.cnl_decode_detections_packed.restype = C.c_int
    lib.cnl_decode_detections_packed.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
        C.c_void_p, C.c_size_t, C.c_void_p]
    lib.cnl_boxes_xyxy_to_xywh.restype = C.c_int
    lib.cnl_boxes_xyxy_to_xywh.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.cnl_sigmoid.restype = C.c_int
    lib.cnl_sigmoid.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.cnl_normalize_images_u8.restype = C.c_int
    lib.cnl_normalize_images_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                            C.POINTER(C.c_float), C.c_void_p]
    lib.cnl_track_workspace_bytes.restype = C.c_size_t
    lib.cnl_track_workspace_bytes.argtypes = [C.c_int, C.c_int]
    lib.cnl_track_cost_matrices.restype = C.c_int
    lib.cnl_track_cost_matrices.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.cnl_gather_boxes.restype = C.c_int
    lib.cnl_gather_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    lib.cnl_engine_create.restype = C.c_int
    lib.cnl_engine_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(BufferDesc), C.c_int,
                                      C.POINTER(ConvDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.cnl_engine_destroy.restype = None
    lib.cnl_engine_destroy.argtypes = [C.c_void_p]
    lib.cnl_engine_arena_bytes.restype = C.c_size_t
    lib.cnl_engine_arena_bytes.argtypes = [C.c_void_p]
    lib.cnl_engine_buffer_offset.restype = C.c_size_t
    lib.cnl_engine_buffer_offset.argtypes = [C.c_void_p, C.c_int]
    lib.cnl_engine_op_form.restype = C.c_int
    lib.cnl_engine_op_form.argtypes = [C.c_void_p, C.c_int]
    lib.cnl_engine_upload.restype = C.c_int
    lib.cnl_engine_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cnl_engine_forward.restype = C.c_int
    lib.cnl_engine_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.POINTER(C.c_int)]
    lib.cnl_engine_forward_act.restype = C.c_int
    lib.cnl_engine_forward_act.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.POINTER(C.c_int)]
    lib.cnl_engine_read_buffer.restype = C.c_int
    lib.cnl_engine_read_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.cnl_engine_write_buffer.restype = C.c_int
    lib.cnl_engine_write_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().cnl_last_error().decode(errors="replace")
        exc = ValueError if status == 1 else (NotImplementedError if status == 2 else CnlError)
        raise exc(f"{what}: {msg} (cnl_status {status})")
