"""CPU: the C-ABI library loads and exports every symbol include/oc_b200.h
declares; host-only entry points behave (no kernel launches here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "oc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from orthogonal_classifiers_b200 import _lib
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in oc_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert _lib.lib.oc_version() >= 100


def test_mps_layout_matches_natural_bond_profile():
    from orthogonal_classifiers_b200.batched import mps_layout
    from orthogonal_classifiers_b200.synthetic import natural_bonds
    for L in (1, 4, 7, 10):
        for D in (None, 1, 2, 3, 10, 16, 32, 100):
            bonds, offs = mps_layout(L, D)
            assert bonds == natural_bonds(L, D)
            assert offs[-1] == sum(2 * a * b for a, b in zip(bonds[:-1], bonds[1:]))
    assert mps_layout(10, 32)[1][-1] == 2728
    assert mps_layout(10, 10)[1][-1] == 888
    assert mps_layout(10, 2)[1][-1] == 72


def test_pack_classifier_layout_and_errors():
    from orthogonal_classifiers_b200 import _lib
    from orthogonal_classifiers_b200.synthetic import random_classifier
    sites = random_classifier(10, 32, centre=5)
    L = len(sites)
    ptrs = (C.c_void_p * L)(*[w.ctypes.data for w in sites])
    shapes = _lib.i32_array([d for w in sites for d in w.shape])
    n = C.c_int64()
    bonds = (C.c_int32 * (L + 1))()
    s = (C.c_int32 * L)()
    assert _lib.lib.oc_pack_classifier(ptrs, shapes, L, 0, 0, None, 0, C.byref(n), bonds, s) == 0
    assert n.value == 18088
    out = np.empty(n.value, np.float32)
    assert _lib.lib.oc_pack_classifier(ptrs, shapes, L, 0, 0, out.ctypes.data, out.size, C.byref(n), bonds, s) == 0
    np.testing.assert_array_equal(out, np.concatenate([w.reshape(-1) for w in sites]).astype(np.float32))
    assert list(s) == [1] * 5 + [16] + [1] * 4 and list(bonds) == [1, 2, 4, 8, 16, 32, 16, 8, 4, 2, 1]
    # capacity error
    assert _lib.lib.oc_pack_classifier(ptrs, shapes, L, 0, 0, out.ctypes.data, 10, C.byref(n), bonds, s) == 4
    # complex128 with zero imaginary part is accepted, non-zero is refused
    csites = [w.astype(np.complex128) for w in sites]
    cptrs = (C.c_void_p * L)(*[w.ctypes.data for w in csites])
    out64 = np.empty(n.value, np.float64)
    assert _lib.lib.oc_pack_classifier(cptrs, shapes, L, 1, 1, out64.ctypes.data, out64.size, C.byref(n), bonds, s) == 0
    np.testing.assert_array_equal(out64, np.concatenate([w.reshape(-1) for w in sites]))
    csites[3][0, 0, 1, 2] += 1e-9j
    assert _lib.lib.oc_pack_classifier(cptrs, shapes, L, 1, 1, out64.ctypes.data, out64.size, C.byref(n), bonds, s) == 2
    assert b"imaginary" in _lib.lib.oc_last_error()


def test_argument_errors_without_gpu():
    from orthogonal_classifiers_b200 import _lib
    lib = _lib.lib
    assert lib.oc_mps_layout(11, 4, None, None) == 1
    assert lib.oc_encode_batch(None, 0, 4, 784, 784, 10, 32, 0, 0, None, None, None, None) == 1
    assert lib.oc_encode_batch(1, 0, 4, 784, 784, 9, 32, 0, 0, 1, None, None, None) == 1   # L mismatch
    assert lib.oc_encode_batch(1, 0, 4, 784, 100, 10, 32, 0, 0, 1, None, None, None) == 1  # ld < n_pix
    assert lib.oc_encode_batch(1, 0, 0, 784, 784, 10, 32, 0, 0, 1, None, None, None) == 0  # empty batch
    assert lib.oc_encode_batch(1, 0, 4, 784, 784, 10, 32, 0, 2, 1, None, None, None) == 1  # unknown trunc_mode
    assert lib.oc_encode_batch(1, 0, -1, 784, 784, 10, 32, 0, 1, 1, None, None, None) == 1  # N < 0
    assert lib.oc_encode_batch(1, 0, 0, 784, 784, 10, 10, 0, 1, 1, None, None, None) == 0   # OC_TRUNC_SWEEP, empty
    # oc_encode_classify / oc_classify_host validate before touching anything
    assert lib.oc_encode_classify(1, 0, -5, 784, 784, 10, 32, 0, 0, 1, None, None, None, 1, None, None) == 1
    assert lib.oc_encode_classify(1, 9, 5, 784, 784, 10, 32, 0, 0, 1, None, None, None, 1, None, None) == 1
    assert lib.oc_encode_classify(1, 0, 5, 784, 784, 10, 32, 5, 0, 1, None, None, None, 1, None, None) == 1
    assert lib.oc_classify_host(1, 0, -5, 784, 784, 10, 32, 0, 0, 1, 10, None, None, 1, None, None) == 1
    assert lib.oc_classify_host(1, 0, 5, 784, 784, 77, 32, 0, 0, 1, 10, None, None, 1, None, None) == 1
    assert lib.oc_set_option(b"nope", 1) == 1
    with pytest.raises(_lib.OcError):
        _lib.check(lib.oc_set_option(b"nope", 1))


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "orthogonal_classifiers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
