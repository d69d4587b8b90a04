"""CPU tests of the drop-in boundary: libudh.so builds/loads, exports every symbol include/udh.h declares, keeps
the Python and C parameter layouts identical, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "udh.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(udh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from unsuperviseddeephomographyral2018_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 19
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), "libudh.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and include/udh.h disagree"
    assert _lib.lib.udh_version() >= 100
    assert isinstance(_lib.lib.udh_last_error(), bytes)


def test_parameter_layout_matches_between_python_and_c():
    from unsuperviseddeephomographyral2018_b200 import _lib, params
    specs = params.param_specs(128)
    assert params.num_parameters(specs) == 34192264
    assert _lib.lib.udh_param_total_floats(128) == params.total_floats(specs)
    for i, s in enumerate(specs.values()):
        off, n = ctypes.c_size_t(), ctypes.c_size_t()
        assert _lib.lib.udh_param_offset(128, i, ctypes.byref(off), ctypes.byref(n)) == 0
        assert (off.value, n.value) == (s.offset, s.size), s.name
        assert s.offset % 32 == 0
    assert _lib.lib.udh_param_offset(128, 20, ctypes.byref(off), ctypes.byref(n)) == _lib.EINVAL
    assert b"bad tensor index" in _lib.lib.udh_last_error()
    assert _lib.lib.udh_cnn_workspace_bytes(128, 128, 0) > 128 * 13_000_000   # ~13.2 MB of activations per pair
    assert _lib.lib.udh_cnn_workspace_bytes(4, 100, 0) == 0                   # unsupported patch size


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_gpu():
    from unsuperviseddeephomographyral2018_b200 import _lib, engine, ops
    assert _lib.lib.udh_device_available() == 0
    with pytest.raises(_lib.UdhError):
        _lib.require_device()
    with pytest.raises(_lib.UdhError):
        engine.HomographyEngine(2)
    with pytest.raises(_lib.UdhError):
        ops.dlt_forward(torch.zeros(1, 8), torch.zeros(1, 8))


def test_schedule_matches_reference_constants():
    from unsuperviseddeephomographyral2018_b200 import engine
    assert engine.decay_steps(1e-4, 0.9e-4) == 58117 and engine.decay_steps(5e-4, 0.9e-4) == 3570
    assert engine.learning_rate(58117, 1e-4, 0.9e-4) == pytest.approx(0.96e-4)
