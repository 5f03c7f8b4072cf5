"""MXNet checkpoint formats (yolo_b200/mxnet_io.py): the NDArray list binary in its v2 / v1 / legacy layouts, the gluon parameter
naming of the reference's constructors and the export symbol JSON - round trips on files written here in those formats (no MXNet
offline)."""
import struct

import numpy as np
import pytest

from oracle import nets
from yolo_b200 import mxnet_io as mio


def _rand_params(spec, net="carnet", seed=0):
    rng = np.random.default_rng(seed)
    return {n: rng.standard_normal(s).astype(np.float32) for n, s in nets.param_shapes(net, spec)}


@pytest.mark.parametrize("version", [2, 1, 0])
def test_ndarray_list_roundtrip(version, tmp_path):
    rng = np.random.default_rng(1)
    d = {"carnet0_conv0_weight": rng.standard_normal((8, 3, 3, 3)).astype(np.float32), "carnet0_batchnorm0_gamma": rng.standard_normal(8).astype(np.float32),
         "h": rng.standard_normal((2, 5)).astype(np.float16), "i": np.arange(6, dtype=np.int32).reshape(3, 2), "u": np.arange(4, dtype=np.uint8)}
    p = tmp_path / "x.params"
    mio.save_ndarray_file(str(p), d, version=version)
    names, arrays = mio.load_ndarray_file(str(p))
    assert names == list(d)
    for a, (k, v) in zip(arrays, d.items()):
        assert a.dtype == v.dtype and a.shape == v.shape and np.array_equal(a, v), k


def test_known_bytes_v2():
    """A hand-assembled file: one float32 (2,3) array named 'w' - byte for byte what NDArray::Save writes (magic, stype, shape as
    uint32 ndim + int64 dims, cpu context, type flag 0, raw data)."""
    data = np.arange(6, dtype="<f4")
    blob = struct.pack("<QQQ", 0x112, 0, 1) + struct.pack("<Ii", 0xF993FAC9, 0) + struct.pack("<Iqq", 2, 2, 3) + struct.pack("<iii", 1, 0, 0) + \
        data.tobytes() + struct.pack("<Q", 1) + struct.pack("<Q", 1) + b"w"
    assert mio.save_ndarray_file(None, {"w": data.reshape(2, 3)}) == blob
    d = mio.load_params_dict(blob)
    assert list(d) == ["w"] and np.array_equal(d["w"], data.reshape(2, 3))
    with pytest.raises(ValueError):
        mio.load_ndarray_file(blob[:-3])                                        # truncated
    with pytest.raises(ValueError):
        mio.load_ndarray_file(struct.pack("<QQQ", 0x113, 0, 0))                 # wrong list magic
    sparse = struct.pack("<QQQ", 0x112, 0, 1) + struct.pack("<Ii", 0xF993FAC9, 1)
    with pytest.raises(ValueError):
        mio.load_ndarray_file(sparse + b"\0" * 64)


def test_gluon_names_of_a_fresh_process():
    """Spot checks of the creation-order naming (yolo_modules/basic_yolo.py:16-39,108-123)."""
    spec = nets.spec_dk53()
    shapes = dict(nets.param_shapes("carnet", spec))
    g = mio.gluon_names("carnet", spec, shapes)
    assert g["stages.0.weight"] == "carnet0_conv0_weight" and g["stages.0.running_var"] == "carnet0_batchnorm0_running_var"
    assert g["stages.1.0.weight"] == "carnet0_conv1_weight" and g["stages.1.1.body.0.weight"] == "carnet0_conv2_weight"
    assert g["stages.5.4.body.1.gamma"] == "carnet0_batchnorm51_gamma"                       # 52 backbone convs
    assert g["yolo_outputs.0.weight"] == "conv0_weight" and g["yolo_outputs.0.bias"] == "conv0_bias"
    assert g["yolo_blocks.0.body.0.weight"] == "yolodetectionblockv30_conv0_weight" and g["yolo_blocks.0.tip.beta"] == "yolodetectionblockv30_batchnorm5_beta"
    assert g["yolo_outputs.1.weight"] == "conv1_weight" and g["transitions.0.weight"] == "conv2_weight" and g["transitions.0.gamma"] == "batchnorm0_gamma"
    assert g["yolo_outputs.2.weight"] == "conv3_weight" and g["transitions.1.weight"] == "conv4_weight" and g["transitions.1.beta"] == "batchnorm1_beta"
    assert len(set(g.values())) == len(g) == len(shapes)


@pytest.mark.parametrize("net,export,offsets", [("carnet", False, None), ("carnet", True, {"conv": 7, "batchnorm": 3, "yolodetectionblockv3": 2}),
                                                ("carlpnet", False, None)])
def test_gluon_checkpoint_roundtrip(net, export, offsets, tmp_path):
    """canonical -> gluon-named file (shuffled like a python-2 dict, global counters offset like a process that built other nets first)
    -> canonical: identical."""
    spec = nets.spec_tiny(size=(64, 96), C=9, lp=(net == "carlpnet"))
    params = _rand_params(spec, net)
    shapes = {k: v.shape for k, v in params.items()}
    names = mio.gluon_names(net, spec, shapes, net_prefix="carlpnet0_" if net == "carlpnet" else "carnet0_", counters=offsets, export=export)
    order = list(params)
    np.random.default_rng(5).shuffle(order)
    p = tmp_path / "export-0000.params"
    mio.save_ndarray_file(str(p), {names[k]: params[k] for k in order})
    back = mio.load_gluon_params(str(p), net, spec, nets.param_shapes(net, spec))
    assert set(back) == set(params)
    for k in params:
        assert np.array_equal(back[k], params[k]), k
    # a missing parameter and a wrong shape are errors (the reference falls back to Xavier init on a failed load)
    broken = {names[k]: params[k] for k in order if k != "stages.1.0.gamma"}
    with pytest.raises((KeyError, ValueError)):
        mio.load_gluon_params(mio.save_ndarray_file(None, broken), net, spec, nets.param_shapes(net, spec))
    bad = {names[k]: (params[k] if k != "yolo_outputs.0.bias" else np.zeros(5, np.float32)) for k in order}
    with pytest.raises(ValueError):
        mio.load_gluon_params(mio.save_ndarray_file(None, bad), net, spec, nets.param_shapes(net, spec))


def test_save_gluon_params_is_loadable(tmp_path):
    spec = nets.spec_micro()
    params = _rand_params(spec)
    p = tmp_path / "w.params"
    mio.save_gluon_params(str(p), params, "carnet", spec)
    back = mio.load_gluon_params(str(p), "carnet", spec, nets.param_shapes("carnet", spec))
    assert all(np.array_equal(back[k], params[k]) for k in params)


@pytest.mark.parametrize("spec", [nets.spec_dk53(), nets.spec_v1_native(), nets.spec_tiny()])
def test_symbol_json_roundtrip(spec, tmp_path):
    shapes = dict(nets.param_shapes("carnet", spec))
    p = tmp_path / "export-symbol.json"
    mio.write_symbol_json(str(p), "carnet", spec, shapes)
    ops = mio.read_symbol_json(str(p))
    got = mio.spec_from_symbol(ops, size=spec["size"])
    assert got["layers"] == spec["layers"] and got["channels"] == spec["channels"]
    A, C = len(spec["all_anchors"][0]), spec["slice_point"][-1]
    assert got["n_scales"] == len(spec["all_anchors"]) and got["head_channels"] == [A * C] * got["n_scales"]
    kinds = {o["op"] for o in ops}
    assert {"Convolution", "BatchNorm", "LeakyReLU", "elemwise_add", "UpSampling", "Concat", "transpose", "Reshape"} <= kinds
    # variables of the graph are exactly the gluon parameter names (+ data)
    import json
    g = json.loads(open(p).read())
    variables = {g["nodes"][i]["name"] for i in g["arg_nodes"]}
    assert variables == set(mio.gluon_names("carnet", spec, shapes).values()) | {"data"}
