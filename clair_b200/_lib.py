"""ctypes binding of libclair_b200.so (the C-ABI in include/clair_b200.h).

There is no CPU fallback: if the shared library is missing this raises, and clairb_create
itself refuses anything that is not an sm_100 device.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLAIRB_LIB") or os.path.join(_HERE, "lib", "libclair_b200.so")   # CLAIRB_LIB: A/B builds
# the cross-check build (-DCLAIRB_CROSSCHECK: tensor-core engine + fp32 CUDA-core engines selected by CLAIRB_ENGINE /
# CLAIRB_FUSED_TAIL / CLAIRB_L2_STREAM); only tests load it: Clair(library=_lib.XCHECK_PATH)
XCHECK_PATH = os.path.join(_HERE, "lib", "libclair_b200_xcheck.so")

OK, EINVAL, ECUDA, ENOMEM, ENODEVICE, EWEIGHTS = range(6)
DTYPE_F32, DTYPE_I16 = 0, 1
N_OUT = 90
DECISION_WORDS = 8
SITE_ELEMS = 1056
LAYER_LSTM1, LAYER_LSTM2, LAYER_L3, LAYER_L4, LAYER_LOGITS = 1, 2, 3, 4, 5

# every symbol include/clair_b200.h declares: (restype, argtypes)
_c = ctypes
SYMBOLS = {
    "clairb_create": (_c.c_int, [_c.c_int, _c.c_int64, _c.c_int, _c.POINTER(_c.c_void_p)]),
    "clairb_set_weight": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_void_p, _c.POINTER(_c.c_int64), _c.c_int]),
    "clairb_finalize_weights": (_c.c_int, [_c.c_void_p]),
    "clairb_predict": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p]),
    "clairb_predict_split": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "clairb_predict_async": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                        _c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_int64)]),
    "clairb_predict_wait": (_c.c_int, [_c.c_void_p, _c.c_int64]),
    "clairb_predict_async_to_device": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p, _c.POINTER(_c.c_int64)]),
    "clairb_predict_to_device": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p]),
    "clairb_predict_device": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p, _c.c_void_p]),
    "clairb_predict_decide": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "clairb_predict_split_decide": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                               _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "clairb_decide": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p]),
    "clairb_create_tensors": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "clairb_predict_created": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p]),
    "clairb_encode_sam": (_c.c_int, [_c.c_char_p, _c.c_int64, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_int64,
                                     _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                     _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int64)]),
    "clairb_blosc_decompress": (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int64, _c.POINTER(_c.c_int64)]),
    "clairb_format_tensor_rows": (_c.c_int, [_c.c_char_p, _c.c_void_p, _c.c_char_p, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_int64,
                                             _c.c_void_p, _c.c_int64, _c.POINTER(_c.c_int64)]),
    "clairb_format_vcf_rows": (_c.c_int, [_c.c_int64, _c.c_char_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                          _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.POINTER(_c.c_int64), _c.c_void_p]),
    "clairb_decode_rows": (_c.c_int, [_c.c_char_p, _c.c_int64, _c.c_int64, _c.c_int, _c.c_void_p, _c.c_void_p,
                                      _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int64)]),
    "clairb_get_layer": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int64]),
    "clairb_host_alloc": (_c.c_int, [_c.POINTER(_c.c_void_p), _c.c_int64]),
    "clairb_host_free": (_c.c_int, [_c.c_void_p]),
    "clairb_kernel_launches": (_c.c_int64, [_c.c_void_p]),
    "clairb_set_profiling": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "clairb_read_profile": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_int64]),
    "clairb_trainer_create": (_c.c_int, [_c.c_int, _c.c_int64, _c.POINTER(_c.c_void_p)]),
    "clairb_trainer_set_weight": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_void_p, _c.POINTER(_c.c_int64), _c.c_int]),
    "clairb_trainer_get": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_char_p, _c.c_void_p, _c.c_int64]),
    "clairb_trainer_set_grad_buffer": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "clairb_trainer_set_dropout_rates": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "clairb_trainer_num_params": (_c.c_int64, [_c.c_void_p]),
    "clairb_trainer_dense_offset": (_c.c_int64, [_c.c_void_p]),
    "clairb_trainer_forward_backward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_uint64,
                                                   _c.c_void_p]),
    "clairb_trainer_backward_lstm": (_c.c_int, [_c.c_void_p]),
    "clairb_trainer_apply": (_c.c_int, [_c.c_void_p, _c.c_float, _c.c_float, _c.c_float, _c.c_int64, _c.POINTER(_c.c_double)]),
    "clairb_trainer_stream": (_c.c_void_p, [_c.c_void_p]),
    "clairb_trainer_set_deferred": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p]),
    "clairb_trainer_read_losses": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_double)]),
    "clairb_trainer_step": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_uint64, _c.c_float,
                                       _c.c_float, _c.c_float, _c.c_int64, _c.POINTER(_c.c_double), _c.POINTER(_c.c_double)]),
    "clairb_trainer_get_probabilities": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64]),
    "clairb_trainer_kernel_launches": (_c.c_int64, [_c.c_void_p]),
    "clairb_trainer_last_error": (_c.c_char_p, [_c.c_void_p]),
    "clairb_trainer_destroy": (_c.c_int, [_c.c_void_p]),
    "clairb_version": (_c.c_char_p, []),
    "clairb_last_error": (_c.c_char_p, [_c.c_void_p]),
    "clairb_destroy": (_c.c_int, [_c.c_void_p]),
}

_libs = {}


def load(path=None):
    """Load the shared library (default: the product build) and attach prototypes.  Raises if it was never built."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(
            "clair_b200: %s not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _libs[path] = lib
    return lib


def last_error(handle=None, lib=None):
    msg = (lib or load()).clairb_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, handle=None, what="", lib=None):
    if rc == OK:
        return
    msg = "%s: %s" % (what, last_error(handle, lib)) if what else last_error(handle, lib)
    if rc in (EINVAL, EWEIGHTS):
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
