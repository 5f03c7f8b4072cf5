"""Reader / writer for the reference's checkpoint files, so that released MXNet weights can be loaded into ``yolo_b200``.

The reference stores weights in two ways (yolo_modules/yolo_gluon.py, paths under the reference repo):

* ``net.collect_params().save(path)`` / ``.load(path)``  (``init_NN`` :172-201, checkpoints in car/YOLO.py:546-549) - an MXNet
  *NDArray list file* whose names are the gluon parameter names (``carnet0_conv0_weight``, ``carnet0_batchnorm0_running_mean`` ...)
* ``net.export(folder + '/export')``  (``export`` :245-262, loaded by ``init_executor`` :204-242 through
  ``mxnet.model.load_checkpoint``) - ``export-symbol.json`` (the graph) + ``export-0000.params`` with ``arg:`` / ``aux:`` prefixed names.

MXNet itself is not installable here, so both formats are restated from MXNet's sources:

NDArray list file (``src/ndarray/ndarray.cc``: ``NDArray::Save`` / ``Load``, ``kMXAPINDArrayListMagic``), little endian::

    uint64 0x112 | uint64 0 | uint64 n_arrays | n_arrays x NDArray | uint64 n_names | n_names x (uint64 len | bytes)
    NDArray  v2/v3: uint32 magic (0xF993FAC9 / 0xF993FACA) | int32 storage_type (0 = dense) | shape | int32 dev_type | int32 dev_id
                    | int32 type_flag | raw data
             v1   : uint32 0xF993FAC8 | shape | ctx | type_flag | data
             legacy: uint32 ndim | ndim x uint32 | ctx | type_flag | data      (no magic: the first word IS ndim)
    shape    v1/v2: uint32 ndim | ndim x int64        v3: int32 ndim | ndim x int64
    type_flag: 0 float32, 1 float64, 2 float16, 3 uint8, 4 int32, 5 int8, 6 int64

Symbol JSON (``nnvm`` graph): ``{"nodes": [{"op", "name", "attrs"|"attr"|"param", "inputs": [[node, out, ver], ...]}], "arg_nodes", "heads"}``.

Gluon parameter names depend on creation-order counters (``conv12``) and on which name scope was current when a block was created;
the mapping below follows gluon's rules for the reference's constructors (yolo_modules/basic_yolo.py:8-39,108-123; gluoncv
``_conv2d`` / ``DarknetBasicBlockV3`` / ``YOLODetectionBlockV3``) and is robust to counter offsets: parameters are matched by
(scope, block type, RANK of the counter, leaf), never by the absolute counter value.  No real checkpoint exists offline, so this is
pinned by round-trip tests on files written here (tests/test_mxnet_io.py), not against MXNet output.
"""
from __future__ import annotations

import json
import re
import struct

import numpy as np

LIST_MAGIC = 0x112
V1_MAGIC, V2_MAGIC, V3_MAGIC = 0xF993FAC8, 0xF993FAC9, 0xF993FACA
DTYPES = {0: np.float32, 1: np.float64, 2: np.float16, 3: np.uint8, 4: np.int32, 5: np.int8, 6: np.int64}
TYPE_FLAGS = {np.dtype(v): k for k, v in DTYPES.items()}


class _Reader:
    def __init__(self, data):
        self.d, self.o = memoryview(data), 0

    def take(self, fmt):
        n = struct.calcsize(fmt)
        if self.o + n > len(self.d):
            raise ValueError("truncated MXNet NDArray file")
        v = struct.unpack_from(fmt, self.d, self.o)
        self.o += n
        return v if len(v) > 1 else v[0]

    def raw(self, n):
        if self.o + n > len(self.d):
            raise ValueError("truncated MXNet NDArray file")
        b = self.d[self.o:self.o + n]
        self.o += n
        return b


def _read_ndarray(r):
    magic = r.take("<I")
    if magic in (V2_MAGIC, V3_MAGIC):
        stype = r.take("<i")
        if stype != 0:
            raise ValueError(f"sparse NDArray storage type {stype} is not supported (dense parameters only)")
        ndim = r.take("<i" if magic == V3_MAGIC else "<I")
        shape = [r.take("<q") for _ in range(max(ndim, 0))]
    elif magic == V1_MAGIC:
        ndim = r.take("<I")
        shape = [r.take("<q") for _ in range(ndim)]
    else:                                  # legacy: the word just read is ndim, dims are uint32
        ndim = magic
        if ndim > 32:
            raise ValueError(f"not an MXNet NDArray (leading word 0x{magic:x})")
        shape = [r.take("<I") for _ in range(ndim)]
    if ndim == 0 and magic != V3_MAGIC:
        return np.zeros((0,), np.float32)   # "none" array: nothing else was written
    r.take("<ii")                           # context (dev_type, dev_id): irrelevant on load
    flag = r.take("<i")
    if flag not in DTYPES:
        raise ValueError(f"unknown MXNet type flag {flag}")
    dt = np.dtype(DTYPES[flag]).newbyteorder("<")
    n = int(np.prod(shape)) if shape else 1
    a = np.frombuffer(r.raw(n * dt.itemsize), dtype=dt).reshape(shape)
    return a.astype(DTYPES[flag], copy=True)


def load_ndarray_file(path_or_bytes):
    """-> (names or None, [np.ndarray]).  ``mx.nd.load`` for dense arrays."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else open(path_or_bytes, "rb").read()
    r = _Reader(data)
    magic, _reserved = r.take("<QQ")
    if magic != LIST_MAGIC:
        raise ValueError(f"not an MXNet NDArray list file (header 0x{magic:x}, expected 0x{LIST_MAGIC:x})")
    arrays = [_read_ndarray(r) for _ in range(r.take("<Q"))]
    n_names = r.take("<Q")
    names = [bytes(r.raw(r.take("<Q"))).decode("utf-8") for _ in range(n_names)]
    if names and len(names) != len(arrays):
        raise ValueError(f"{len(names)} names for {len(arrays)} arrays")
    return (names or None), arrays


def load_params_dict(path_or_bytes):
    """name -> array (``mx.nd.load`` of a dict file)."""
    names, arrays = load_ndarray_file(path_or_bytes)
    if names is None:
        raise ValueError("the file holds a list, not a dict (no names)")
    return dict(zip(names, arrays))


def save_ndarray_file(path, params, version=2):
    """``mx.nd.save(path, dict)``: writes the v2 layout (what MXNet 1.x reads and writes); ``version=1`` / ``0`` (legacy) for tests."""
    out = [struct.pack("<QQQ", LIST_MAGIC, 0, len(params))]
    for a in params.values():
        a = np.ascontiguousarray(a)
        if a.dtype not in TYPE_FLAGS:
            raise ValueError(f"dtype {a.dtype} has no MXNet type flag")
        if version >= 2:
            out.append(struct.pack("<Ii", V2_MAGIC, 0))
            out.append(struct.pack("<I", a.ndim) + b"".join(struct.pack("<q", s) for s in a.shape))
        elif version == 1:
            out.append(struct.pack("<I", V1_MAGIC))
            out.append(struct.pack("<I", a.ndim) + b"".join(struct.pack("<q", s) for s in a.shape))
        else:
            out.append(struct.pack("<I", a.ndim) + b"".join(struct.pack("<I", s) for s in a.shape))
        out.append(struct.pack("<iii", 1, 0, TYPE_FLAGS[a.dtype]))          # cpu(0), type flag
        out.append(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())
    out.append(struct.pack("<Q", len(params)))
    for k in params:
        b = k.encode("utf-8")
        out.append(struct.pack("<Q", len(b)) + b)
    blob = b"".join(out)
    if path is not None:
        with open(path, "wb") as f:
            f.write(blob)
    return blob


# ---------------------------------------------------------------------------------------------------------------------------
# gluon parameter names  <->  canonical names of the C library (net.param_shapes())
# ---------------------------------------------------------------------------------------------------------------------------
_LEAVES = {"weight": "weight", "bias": "bias", "gamma": "gamma", "beta": "beta", "running_mean": "running_mean", "running_var": "running_var",
           "moving_mean": "running_mean", "moving_var": "running_var"}
_NAME_RE = re.compile(r"^(?P<scope>.*?)(?P<type>conv|batchnorm|dense)(?P<idx>\d+)_(?P<leaf>weight|bias|gamma|beta|running_mean|running_var|moving_mean|moving_var)$")


def gluon_layout(net_type, spec):
    """The convolutions of the reference's network in gluon CREATION order, grouped by the name scope their parameters land in.

    Returns a list of groups ``(scope_kind, scope_rank, [canonical conv names in creation order])``:
      ("net", 0, ...)        created inside ``with self.name_scope()`` of BasicYOLONet: stem, stage down-convs, DarknetBasicBlockV3 bodies
                             (that block opens no name scope of its own) -> ``<net prefix>conv<i>_``, ``<net prefix>batchnorm<i>_``
      ("block", j, ...)      YOLODetectionBlockV3 number j (opens its own scope)  -> ``yolodetectionblockv3<j>_conv<i>_``
      ("outer", 0, ...)      created outside any scope by YOLOPyrmaid: YOLOOutput convs (bias, no BN) and transition ``_conv2d`` cells,
                             interleaved in creation order                        -> ``conv<i>_`` / ``batchnorm<i>_`` (global counters)
    CarLPNet adds its LP branch blocks the same way (car_and_LP/YOLO.py:47-61)."""
    if net_type not in ("carnet", "carlpnet"):
        raise ValueError("gluon name mapping is implemented for the YOLO topologies (carnet / carlpnet)")
    layers = spec["layers"]
    nl, npyr = len(layers), len(spec["all_anchors"])
    net = ["stages.0"]
    for st in range(1, nl + 1):
        net.append(f"stages.{st}.0")
        for j in range(1, layers[st - 1] + 1):
            net += [f"stages.{st}.{j}.body.0", f"stages.{st}.{j}.body.1"]
    groups = [("net", 0, net)]
    outer = []
    for i in range(npyr):
        outer.append(f"yolo_outputs.{i}")
        groups.append(("block", i, [f"yolo_blocks.{i}.body.{k}" for k in range(5)] + [f"yolo_blocks.{i}.tip"]))
        if i > 0:
            outer.append(f"transitions.{i - 1}")
    if net_type == "carlpnet":
        for b in range(5):
            groups.append(("block", npyr + b, [f"LP_branch.{b}.body.{k}" for k in range(5)] + [f"LP_branch.{b}.tip"]))
        outer.append("LP_branch.5")
    groups.append(("outer", 0, outer))
    return groups


def map_gluon_names(names, net_type, spec, shapes):
    """gluon name -> canonical name for every entry of ``names`` (``arg:`` / ``aux:`` prefixes accepted).

    ``shapes``: canonical name -> shape (``dict(net.param_shapes())``).  Matching is by (scope, type, rank of the counter within that
    scope and type, leaf); every match is shape-checked by the caller."""
    has_bn = {n.rsplit(".", 1)[0] for n in shapes if n.endswith(".gamma")}
    parsed = {}
    for full in names:
        n = full.split(":", 1)[1] if full[:4] in ("arg:", "aux:") else full
        m = _NAME_RE.match(n)
        if not m:
            raise ValueError(f"'{full}' does not look like a gluon conv / batchnorm parameter name")
        parsed[full] = (m["scope"], m["type"], int(m["idx"]), _LEAVES[m["leaf"]])
    # scopes: the one containing 'yolodetectionblockv3' -> blocks (ranked by their counter); the net scope holds the most convs
    scope_conv_count = {}
    for scope, typ, idx, leaf in parsed.values():
        if typ == "conv" and leaf == "weight":
            scope_conv_count[scope] = scope_conv_count.get(scope, 0) + 1
    block_scopes = sorted((s for s in scope_conv_count if "yolodetectionblockv3" in s),
                          key=lambda s: int(re.search(r"yolodetectionblockv3(\d+)_", s).group(1)))
    rest = [s for s in scope_conv_count if s not in block_scopes]
    groups = gluon_layout(net_type, spec)
    n_net = len(groups[0][2])
    net_scope = [s for s in rest if scope_conv_count[s] == n_net]
    outer_scope = [s for s in rest if s not in net_scope]
    if len(net_scope) != 1 or len(outer_scope) != 1:
        raise ValueError(f"cannot identify the backbone / pyramid name scopes among {sorted(scope_conv_count.items())}")
    scope_of = {("net", 0): net_scope[0], ("outer", 0): outer_scope[0]}
    blocks = [g for g in groups if g[0] == "block"]
    if len(blocks) != len(block_scopes):
        raise ValueError(f"{len(block_scopes)} detection-block scopes in the file, the spec needs {len(blocks)}")
    for g, s in zip(blocks, block_scopes):
        scope_of[(g[0], g[1])] = s
    out = {}
    for kind, rank, convs in groups:
        scope = scope_of[(kind, rank)]
        conv_idx = sorted({idx for (s, t, idx, leaf) in parsed.values() if s == scope and t == "conv"})
        bn_idx = sorted({idx for (s, t, idx, leaf) in parsed.values() if s == scope and t == "batchnorm"})
        if len(conv_idx) != len(convs):
            raise ValueError(f"scope '{scope}': {len(conv_idx)} convolutions in the file, {len(convs)} expected")
        conv_of = dict(zip(conv_idx, convs))
        bn_convs = [c for c in convs if c in has_bn]
        if len(bn_idx) != len(bn_convs):
            raise ValueError(f"scope '{scope}': {len(bn_idx)} batchnorms in the file, {len(bn_convs)} expected")
        bn_of = dict(zip(bn_idx, bn_convs))
        for full, (s, t, idx, leaf) in parsed.items():
            if s != scope:
                continue
            out[full] = f"{conv_of[idx] if t == 'conv' else bn_of[idx]}.{leaf}"
    return out


def gluon_names(net_type, spec, shapes, net_prefix="carnet0_", counters=None, export=False):
    """The gluon names the reference's constructor produces in a fresh process (inverse of map_gluon_names), canonical -> gluon.
    ``counters``: starting values of the global ``conv`` / ``batchnorm`` / ``yolodetectionblockv3`` counters (0 in a fresh process)."""
    counters = dict({"conv": 0, "batchnorm": 0, "yolodetectionblockv3": 0}, **(counters or {}))
    has_bn = {n.rsplit(".", 1)[0] for n in shapes if n.endswith(".gamma")}
    has_bias = {n.rsplit(".", 1)[0] for n in shapes if n.endswith(".bias")}
    out = {}
    for kind, rank, convs in gluon_layout(net_type, spec):
        if kind == "net":
            scope, ci, bi = net_prefix, 0, 0
        elif kind == "block":
            scope, ci, bi = f"yolodetectionblockv3{counters['yolodetectionblockv3'] + rank}_", 0, 0
        else:
            scope, ci, bi = "", counters["conv"], counters["batchnorm"]
        for c in convs:
            out[f"{c}.weight"] = f"{scope}conv{ci}_weight"
            if c in has_bias:
                out[f"{c}.bias"] = f"{scope}conv{ci}_bias"
            ci += 1
            if c in has_bn:
                for leaf in ("gamma", "beta", "running_mean", "running_var"):
                    out[f"{c}.{leaf}"] = f"{scope}batchnorm{bi}_{leaf}"
                bi += 1
    if export:
        out = {k: ("aux:" if k.endswith(("running_mean", "running_var")) else "arg:") + v for k, v in out.items()}
    return out


def load_gluon_params(path_or_bytes, net_type, spec, param_shapes):
    """Read a ``collect_params().save`` / ``export-0000.params`` file into a canonical-name dict ready for ``Net.load_params``.
    ``param_shapes``: ``net.param_shapes()``.  Raises on any missing parameter or shape mismatch (the reference falls back to Xavier
    initialisation when ``load`` throws, yolo_gluon.py:194-198 - the caller decides)."""
    shapes = dict(param_shapes)
    raw = load_params_dict(path_or_bytes)
    mapping = map_gluon_names(list(raw), net_type, spec, shapes)
    out = {}
    for full, canon in mapping.items():
        if canon not in shapes:
            raise KeyError(f"'{full}' maps to '{canon}', which the network does not have")
        a = np.asarray(raw[full], np.float32)
        if tuple(a.shape) != tuple(shapes[canon]):
            raise ValueError(f"'{full}' -> '{canon}': shape {tuple(a.shape)} in the file, {tuple(shapes[canon])} expected")
        out[canon] = a
    missing = [n for n in shapes if n not in out]
    if missing:
        raise KeyError(f"{len(missing)} parameters missing in the file, e.g. {missing[:3]}")
    return out


def save_gluon_params(path, params, net_type, spec, net_prefix="carnet0_", export=False):
    """Write canonical parameters as a file the reference's ``collect_params().load`` (or ``load_checkpoint`` with export=True) reads."""
    shapes = {k: tuple(np.shape(v)) for k, v in params.items()}
    names = gluon_names(net_type, spec, shapes, net_prefix, export=export)
    return save_ndarray_file(path, {names[k]: np.asarray(v, np.float32) for k, v in params.items()})


# ---------------------------------------------------------------------------------------------------------------------------
# export-symbol.json
# ---------------------------------------------------------------------------------------------------------------------------
def _attrs(node):
    return node.get("attrs") or node.get("attr") or node.get("param") or {}


def _tuple(s):
    return tuple(int(v) for v in re.findall(r"-?\d+", s))


def read_symbol_json(path_or_text):
    """-> list of op nodes ``{"op", "name", "attrs", "inputs": [node names]}`` in graph (topological) order, variables skipped."""
    text = path_or_text if path_or_text.lstrip().startswith("{") else open(path_or_text).read()
    g = json.loads(text)
    nodes = g["nodes"]
    ops = []
    for n in nodes:
        if n["op"] == "null":
            continue
        ops.append({"op": n["op"], "name": n["name"], "attrs": _attrs(n), "inputs": [nodes[i[0]]["name"] for i in n["inputs"]]})
    return ops


def spec_from_symbol(ops, size=None):
    """Recover the spec fields that define the YOLO topology (``layers``, ``channels``, the heads' channel counts) from the exported
    graph: the first Convolution is the stem; every stride-2 convolution opens a stage; the residual blocks of a stage are the
    elemwise_add nodes up to the next stage (the pyramid has none); head convolutions are the ones with a bias."""
    convs = [o for o in ops if o["op"] == "Convolution"]
    if not convs:
        raise ValueError("no Convolution node in the symbol")
    channels, layers, adds, open_stage = [int(convs[0]["attrs"]["num_filter"])], [], 0, False
    for o in ops:
        if o["op"] in ("elemwise_add", "_Plus", "_plus", "broadcast_add"):
            adds += 1
        elif o["op"] == "Convolution" and o is not convs[0]:
            stride = _tuple(o["attrs"].get("stride", "(1, 1)"))
            if stride and stride[0] == 2:
                if open_stage:
                    layers.append(adds)
                open_stage, adds = True, 0
                channels.append(int(o["attrs"]["num_filter"]))
    if open_stage:
        layers.append(adds)
    heads = [o for o in convs if str(o["attrs"].get("no_bias", "False")).lower() in ("false", "0")]
    out = {"layers": layers, "channels": channels, "n_scales": len(heads), "head_channels": [int(o["attrs"]["num_filter"]) for o in heads]}
    if size is not None:
        out["size"] = list(size)
    return out


def write_symbol_json(path, net_type, spec, shapes, net_prefix="carnet0_"):
    """Emit the exported graph of the reference's network in MXNet's JSON schema (for round-trip tests and for handing trained weights
    back to an MXNet deployment): Convolution / BatchNorm / LeakyReLU / elemwise_add / UpSampling / Concat / transpose / Reshape nodes
    with the gluon parameter names as variables."""
    names = gluon_names(net_type, spec, shapes, net_prefix)
    nodes, arg_nodes = [], []

    def var(name):
        nodes.append({"op": "null", "name": name, "inputs": []})
        arg_nodes.append(len(nodes) - 1)
        return len(nodes) - 1

    def op(kind, name, inputs, **attrs):
        nodes.append({"op": kind, "name": name, "attrs": {k: str(v) for k, v in attrs.items()}, "inputs": [[i, 0, 0] for i in inputs]})
        return len(nodes) - 1

    def conv_bn_leaky(x, canon, k, pad, stride):
        cout = shapes[canon + ".weight"][0]
        w = var(names[canon + ".weight"])
        c = op("Convolution", names[canon + ".weight"][:-7] + "_fwd", [x, w], kernel=(k, k), stride=(stride, stride), pad=(pad, pad), num_filter=cout,
               no_bias=True)
        bn = names[canon + ".gamma"][:-6]
        ins = [c] + [var(names[f"{canon}.{leaf}"]) for leaf in ("gamma", "beta", "running_mean", "running_var")]
        b = op("BatchNorm", bn + "_fwd", ins, eps=1e-5, momentum=0.9, fix_gamma=False)
        return op("LeakyReLU", bn.replace("batchnorm", "leakyrelu") + "_fwd", [b], act_type="leaky", slope=0.1)

    layers, nl, npyr = spec["layers"], len(spec["layers"]), len(spec["all_anchors"])
    x = conv_bn_leaky(var("data"), "stages.0", 3, 1, 1)
    routes = []
    for st in range(1, nl + 1):
        x = conv_bn_leaky(x, f"stages.{st}.0", 3, 1, 2)
        for j in range(1, layers[st - 1] + 1):
            y = conv_bn_leaky(x, f"stages.{st}.{j}.body.0", 1, 0, 1)
            y = conv_bn_leaky(y, f"stages.{st}.{j}.body.1", 3, 1, 1)
            x = op("elemwise_add", f"{net_prefix}darknetbasicblockv3{len(routes)}_{j}__plus", [x, y])
        if st >= nl + 1 - npyr:
            routes.append(x)
    heads = []
    A, C = len(spec["all_anchors"][0]), spec["slice_point"][-1]
    for i in range(npyr):
        for k in range(5):
            x = conv_bn_leaky(x, f"yolo_blocks.{i}.body.{k}", 1 if k % 2 == 0 else 3, 0 if k % 2 == 0 else 1, 1)
        tip = conv_bn_leaky(x, f"yolo_blocks.{i}.tip", 3, 1, 1)
        w, b = var(names[f"yolo_outputs.{i}.weight"]), var(names[f"yolo_outputs.{i}.bias"])
        o = op("Convolution", names[f"yolo_outputs.{i}.weight"][:-7] + "_fwd", [tip, w, b], kernel=(1, 1), num_filter=A * C, no_bias=False)
        o = op("transpose", f"yolooutput{i}_transpose0", [o], axes=(0, 2, 3, 1))
        heads.append(op("Reshape", f"yolooutput{i}_reshape0", [o], shape=(0, -1, A, C)))
        if i == npyr - 1:
            break
        x = conv_bn_leaky(x, f"transitions.{i}", 1, 0, 1)
        x = op("UpSampling", f"upsampling{i}", [x], scale=2, sample_type="nearest")
        x = op("Concat", f"concat{i}", [x, routes[::-1][i + 1]], dim=1, num_args=2)
    g = {"nodes": nodes, "arg_nodes": arg_nodes, "node_row_ptr": list(range(len(nodes) + 1)), "heads": [[h, 0, 0] for h in heads[::-1]],
         "attrs": {"mxnet_version": ["int", 10400]}}
    text = json.dumps(g, indent=2)
    if path is not None:
        with open(path, "w") as f:
            f.write(text)
    return text
