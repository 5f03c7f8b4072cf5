"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/yolo_b200.h
declares, builds the same parameter inventory as the oracle, and reports errors as codes + text
(never aborts) - no compute calls, so no GPU is needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import nets
from yolo_b200 import _lib, api


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "yolo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yolo_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/yolo_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes table and header disagree"
    assert b"sm_100a" in lib.yolo_version()


def _params(lib, h):
    out = []
    for i in range(lib.yolo_param_count(h)):
        name, shp, nd = C.c_char_p(), (C.c_int32 * 4)(), C.c_int32()
        assert lib.yolo_param_info(h, i, C.byref(name), C.byref(shp), C.byref(nd)) == 0
        out.append((name.value.decode(), tuple(shp[k] for k in range(nd.value))))
    return out


@pytest.mark.parametrize("net,spec,gflop", [
    ("carnet", nets.spec_dk53(), 113.263),
    ("carnet", nets.spec_dk53((608, 608)), 241.941),
    ("carnet", nets.spec_v1_native(), 29.773),
    ("carlpnet", nets.spec_dk53((608, 608), 30, True), 469.121),
    ("carlpnet", nets.spec_v1_native(80, True), 55.117),
    ("lpdensenet", nets.spec_lp_v2(), 6.419),
    ("carnet", nets.spec_tiny(), None),
    ("lpdensenet", nets.spec_lp_tiny(), None),
])
def test_plan_matches_oracle_inventory(lib, net, spec, gflop):
    cs = api.make_c_spec(net, spec, "fp32", 2)
    h = C.c_void_p()
    assert lib.yolo_create(C.byref(cs), 0, C.byref(h)) == 0, lib.yolo_last_error(None)
    try:
        assert _params(lib, h) == nets.param_shapes(net, spec)
        if gflop is not None:
            assert abs(lib.yolo_conv_flops_per_image(h) / 1e9 - gflop) < 2e-3
        assert lib.yolo_workspace_bytes(h, 2) > 0
        assert lib.yolo_workspace_bytes(h, 3) == 0               # beyond max_batch
        n_out = lib.yolo_output_count(h)
        assert n_out == {"carnet": 3, "carlpnet": 4, "lpdensenet": 1}[net]
    finally:
        lib.yolo_destroy(h)


def test_output_shapes(lib):
    cs = api.make_c_spec("carlpnet", nets.spec_dk53((608, 608), 30, True), "fp32", 1)
    h = C.c_void_p()
    assert lib.yolo_create(C.byref(cs), 0, C.byref(h)) == 0
    shapes = []
    for i in range(4):
        shp, nd = (C.c_int32 * 4)(), C.c_int32()
        assert lib.yolo_output_shape(h, i, C.byref(shp), C.byref(nd)) == 0
        shapes.append(tuple(shp[k] for k in range(nd.value)))
    lib.yolo_destroy(h)
    assert shapes == [(76 * 76, 3, 30), (38 * 38, 3, 30), (19 * 19, 3, 30), (76, 76, 10)]


def test_errors_are_codes_with_text(lib):
    h = C.c_void_p()
    # 416 is illegal for the 6-stage car/v1 spec (car/utils.py:93 concat mismatch)
    spec = nets.spec_v1_native(); spec["size"] = [416, 416]
    cs = api.make_c_spec("carnet", spec, "fp32", 1)
    rc = lib.yolo_create(C.byref(cs), 0, C.byref(h))
    assert rc == -2 and b"divisible" in lib.yolo_last_error(None)
    cs = api.make_c_spec("carnet", nets.spec_tiny(), "fp32", 0)
    assert lib.yolo_create(C.byref(cs), 0, C.byref(h)) == -1
    cs = api.make_c_spec("carnet", nets.spec_tiny(), "fp32", 1)
    assert lib.yolo_create(C.byref(cs), 0, C.byref(h)) == 0
    w = np.zeros(7, np.float32)
    assert lib.yolo_load_param(h, b"no.such.param", w.ctypes.data_as(C.c_void_p), 7) == -1
    assert b"unknown parameter" in lib.yolo_last_error(h)
    assert lib.yolo_load_param(h, b"stages.0.weight", w.ctypes.data_as(C.c_void_p), 7) == -2
    assert b"expects" in lib.yolo_last_error(h)
    assert lib.yolo_finalize_params(h, None) == -6               # parameters missing -> YOLO_E_STATE
    x = np.zeros(4, np.float32)
    outs = (C.c_void_p * 3)(1, 1, 1)
    assert lib.yolo_forward(h, x.ctypes.data_as(C.c_void_p), 1, 0, outs, None) == -6   # not finalized
    lib.yolo_destroy(h)
    # decode argument validation happens before any launch
    g = api.make_geom(nets.spec_tiny())
    g.n_scales = 7
    heads = (C.c_void_p * 3)(1, 1, 1)
    assert lib.yolo_decode_top1(C.byref(g), heads, 1, C.c_void_p(1), C.c_void_p(1), None) == -1
    g = api.make_geom(nets.spec_tiny())
    p = _lib.NmsParams(0.5, 0.5, 10, 5000)
    assert lib.yolo_decode_nms(C.byref(g), heads, 1, C.byref(p), C.c_void_p(1), C.c_void_p(1), C.c_void_p(1), None) == -1


def test_python_host_rejects_bad_specs():
    spec = nets.spec_tiny(); spec["slice_point"] = [1, 5, 17]
    with pytest.raises(ValueError):
        api.make_c_spec("carnet", spec)
    spec = nets.spec_tiny(); spec["channels"] = spec["channels"][:-1]
    with pytest.raises(ValueError):
        api.make_c_spec("carnet", spec)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import yolo_b200
    with pytest.raises(RuntimeError):
        yolo_b200.Net("carnet", nets.spec_tiny())
    with pytest.raises(RuntimeError):
        yolo_b200.YOLO(spec=dict(nets.spec_tiny(), classes=[0, 1, 2]))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "yolo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "from .. import oracle" not in text and "oracle." not in text.replace("the oracle", ""), f
