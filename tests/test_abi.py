"""The C-ABI boundary without a GPU: the library loads, exports every symbol the header declares,
its host-side helpers agree with the oracle, and compute calls fail loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from movis_b200 import _native
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "movis_b200.h")).read()
    return sorted(set(re.findall(r"MVB_API[^;(]*?\b(mvb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _header_symbols()
    assert len(names) >= 17
    lib = _native.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/movis_b200.h but not exported"
    assert sorted(_native.EXPORTED_SYMBOLS) == names
    assert lib.mvb_abi_version() == 2 and lib.mvb_max_layers_per_launch() >= 16


def test_layer_struct_layout_matches_header():
    dt = _native.LAYER_DTYPE
    offs = {n: dt.fields[n][1] for n in dt.names}
    assert offs == {"src": 0, "src_w": 8, "src_h": 12, "src_stride": 16, "M": 24, "bbox_x": 72, "bbox_y": 76,
                    "bbox_w": 80, "bbox_h": 84, "blend_mode": 88, "flags": 92, "opacity": 96, "alpha_map": 104}
    assert dt.itemsize == 112


def test_host_helpers_agree_with_oracle():
    for r in [0.1, 0.3, 0.5, 0.77, 1.0, 1.5, 2.0, 3.3, 7.7, 10.0, 17.2, 25.0, 60.0]:
        k = _native.blur_ksize(r)
        assert k == O.blur_ksize(r) == 2 * int(np.ceil(max(1, (4 * r - 1) / 2))) + 1
        assert np.array_equal(_native.gaussian_weights_q8(r, k).astype(np.int32), O.gaussian_weights_q8(r, k))
    rng = np.random.default_rng(0)
    cols = rng.integers(0, 256, (500, 3), dtype=np.uint8)
    want = O.rgb2hsv(cols.reshape(1, -1, 3))[0]
    got = np.stack([_native.rgb_to_hsv_u8(c) for c in cols])
    assert np.array_equal(got, want)


def test_chroma_key_bounds_need_no_gpu():
    from movis_b200.contrib.segmentation import ChromaKey
    e = ChromaKey()
    assert e._lower_color.tolist() == [50, 178, 178] and e._upper_color.tolist() == [70, 255, 255]
    for col, rng_ in [((10, 20, 250), (40.0, 0.5, 0.2)), ((255, 255, 255), (360.0, 1.0, 1.0))]:
        e = ChromaKey(col, key_color_range=rng_)
        lo, hi = O.chroma_key_bounds(col, rng_)
        assert np.array_equal(e._lower_color, lo) and np.array_equal(e._upper_color, hi)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import movis_b200 as mv
    lib = _native.lib()
    assert lib.mvb_device_check() < 0 and lib.mvb_last_error()
    buf = np.zeros((4, 4, 4), np.uint8)
    bg = np.zeros(4, np.uint8)
    rc = lib.mvb_composite_stack(buf.ctypes.data, 4, 4, 16, bg.ctypes.data, 0, None, 0, None)
    assert rc in (-2, -3)
    comp = mv.layer.Composition(size=(8, 8), duration=1.0)
    comp.add_layer(mv.layer.Image.from_color((8, 8), "red"))
    with pytest.raises(RuntimeError):
        comp(0.0)
    with pytest.raises(RuntimeError):
        mv.alpha_composite(buf, buf)
    assert comp(2.0) is None  # range check comes first, as in the reference
