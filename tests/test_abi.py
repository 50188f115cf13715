"""The C-ABI library: loads, exports every symbol include/clair_b200.h declares, and refuses to
run without an sm_100 device (no compute calls here - this file runs on the CPU box)."""
import ctypes
import os
import re

import pytest

from clair_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "clair_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clairb_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in ctypes.cast(lib.clairb_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_header_constants_match_binding():
    text = open(os.path.join(ROOT, "include", "clair_b200.h")).read()
    consts = dict(re.findall(r"#define\s+(CLAIRB_[A-Z0-9_]+)\s+(\d+)", text))
    assert int(consts["CLAIRB_N_OUT"]) == _lib.N_OUT == 90
    assert int(consts["CLAIRB_SITE_ELEMS"]) == _lib.SITE_ELEMS == 33 * 8 * 4
    assert int(consts["CLAIRB_DTYPE_I16"]) == _lib.DTYPE_I16
    assert int(consts["CLAIRB_LAYER_LOGITS"]) == _lib.LAYER_LOGITS
    assert int(consts["CLAIRB_ENODEVICE"]) == _lib.ENODEVICE


def test_create_fails_loudly_without_a_b200():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the refusal path is for the CPU box")
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.clairb_create(0, 1000, 1000, ctypes.byref(h))
    assert rc == _lib.ENODEVICE and not h.value
    assert "not available" in _lib.last_error(None) or "compute capability" in _lib.last_error(None)
    with pytest.raises(RuntimeError):
        from clair_b200.model import Clair
        Clair()


def test_trainer_fails_loudly_without_a_b200():
    """The training step has no CPU path either: the handle refuses to exist without an sm_100 device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the refusal path is for the CPU box")
    lib = _lib.load()
    t = ctypes.c_void_p()
    rc = lib.clairb_trainer_create(0, 512, ctypes.byref(t))
    assert rc == _lib.ENODEVICE and not t.value
    msg = lib.clairb_trainer_last_error(None)
    assert msg and (b"not available" in msg or b"compute capability" in msg)
    from clair_b200.train import Trainer
    with pytest.raises(RuntimeError):
        Trainer()
    assert lib.clairb_trainer_stream(None) is None and lib.clairb_trainer_num_params(None) in (0, -1)


def test_bad_arguments_are_rejected_not_crashed():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.clairb_create(0, 0, 1000, ctypes.byref(h)) == _lib.EINVAL
    assert lib.clairb_predict(None, None, 0, 1, None) == _lib.EINVAL
    assert lib.clairb_destroy(None) == _lib.EINVAL
    assert lib.clairb_kernel_launches(None) == 0
