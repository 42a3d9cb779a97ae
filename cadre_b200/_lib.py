"""ctypes binding of libcadre_sm100.so (C ABI declared in include/cadre_b200.h).

The library is the product: if it is missing or a call fails, this module raises — there is no eager /
CPU fallback anywhere in the package.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcadre_sm100.so")


class CadreError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    """Mirror of `cadre_gemm_args` (include/cadre_b200.h)."""
    _fields_ = [
        ("kind", ctypes.c_int32), ("a_mn", ctypes.c_int32), ("b_mn", ctypes.c_int32), ("batch", ctypes.c_int32),
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("block_n", ctypes.c_int32),
        ("A", ctypes.c_void_p), ("B", ctypes.c_void_p),
        ("lda", ctypes.c_int64), ("a_bs", ctypes.c_int64), ("ldb", ctypes.c_int64), ("b_bs", ctypes.c_int64),
        ("out", ctypes.c_void_p), ("ldc", ctypes.c_int64), ("out_bs", ctypes.c_int64),
        ("out_f32", ctypes.c_int32), ("act", ctypes.c_int32),
        ("bias", ctypes.c_void_p), ("bias_bs", ctypes.c_int64),
        ("res", ctypes.c_void_p), ("ldr", ctypes.c_int64), ("res_bs", ctypes.c_int64),
        ("mask", ctypes.c_void_p), ("ldm", ctypes.c_int64), ("mask_bs", ctypes.c_int64),
        ("res_after_act", ctypes.c_int32), ("rows_is_k", ctypes.c_int32),
        ("batch_rows", ctypes.c_void_p),
        ("alpha", ctypes.c_float), ("epi", ctypes.c_int32),
        ("xpart", ctypes.c_void_p), ("c_prev", ctypes.c_void_p), ("c_out", ctypes.c_void_p),
        ("h_out", ctypes.c_void_p), ("gates_out", ctypes.c_void_p),
        ("ldx", ctypes.c_int64), ("x_bs", ctypes.c_int64), ("ldh", ctypes.c_int64), ("h_bs", ctypes.c_int64),
    ]


_lib = None


def lib():
    """Load (once) and return the shared library; raises CadreError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CadreError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / eager fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.cadre_last_error.restype = ctypes.c_char_p
        _lib.cadre_version.restype = ctypes.c_char_p
    return _lib


def enc_dtype():
    """torch dtype of the encoder's 16-bit operands / activations as compiled into the library."""
    import torch
    return torch.float16 if lib().cadre_enc_dtype() == 1 else torch.bfloat16


def check(rc):
    if rc != 0:
        raise CadreError(f"libcadre_sm100 error {rc}: {lib().cadre_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (default: the current device)."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
