"""Dependency-free reader of the reference's model files: the weights and custom metadata libk2b200.so needs, straight from
decoder*.onnx / joiner*.onnx (and, for the encoder_proj tail, encoder*.onnx).

The reference builds its projs from these files through ONNX Runtime sessions and reads `context_size`, `vocab_size`,
`joiner_dim`, `model_type`, ... from their custom metadata map (ref OfflineModel.cs:31-72, :84-118; OnlineModel.cs:37-166;
README.EN.md:63-66). Neither `onnx` nor `onnxruntime` exists in this image, and the drop-in must not need them either: an ONNX
file is a protobuf `ModelProto` whose initialisers are plain `TensorProto`s, so a wire-format walker of ~150 lines is enough.
`csharp/B200Weights.cs` is the same walker in C#.

What is read (field numbers of onnx.proto3):
  ModelProto     7 graph, 14 metadata_props {1 key, 2 value}
  GraphProto     1 node, 5 initializer
  NodeProto      1 input, 2 output, 4 op_type, 5 attribute {1 name, 3 i, 20 type}
  TensorProto    1 dims, 2 data_type, 4 float_data, 5 int32_data, 7 int64_data, 8 name, 9 raw_data

Which initialiser is which weight is decided by the graph, not by names (exporters rename them `onnx::MatMul_123`):
  decoder.onnx   Gather.data -> emb [V,D];  Conv.W -> conv_w [D,4,ctx];  Gemm / MatMul(+Add) -> dec_proj_w [J,D], dec_proj_b [J]
  joiner.onnx    the Gemm / MatMul(+Add) after Tanh -> out_w [V,J], out_b [V]
  encoder.onnx   (optional) the LAST Gemm / MatMul(+Add) whose output width is joiner_dim -> enc_proj_w [J,E], enc_proj_b [J]
A MatMul stores its weight [in,out] and is transposed to the [out,in] layout of k2b_load_weights; a Gemm honours transB.
Dynamically quantised exports (`*.int8.onnx`: MatMulInteger with `<name>_quantized` / `_scale` / `_zero_point` initialisers, the
onnxruntime quantiser's naming) are de-quantised to fp32.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 10: np.float16, 11: np.float64}


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yields (field number, wire type, value) of one message: varint -> int, 64-bit / 32-bit -> bytes, length-delimited -> bytes."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, v


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [v]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


def _signed(x: int) -> int:
    return x - (1 << 64) if x >= (1 << 63) else x


@dataclass
class Tensor:
    name: str
    array: np.ndarray


@dataclass
class Node:
    op_type: str
    inputs: List[str]
    outputs: List[str]
    ints: Dict[str, int] = field(default_factory=dict)


@dataclass
class OnnxModel:
    metadata: Dict[str, str]
    initializers: Dict[str, np.ndarray]
    nodes: List[Node]


def _tensor(buf: bytes) -> Tensor:
    dims: List[int] = []
    dtype, name, raw = 1, "", None
    floats: List[bytes] = []
    ints: List[int] = []
    for num, wt, v in _fields(buf):
        if num == 1:
            dims += [_signed(x) for x in _packed_varints(v, wt)]
        elif num == 2:
            dtype = v
        elif num == 4:
            floats.append(v)                                   # packed (wt 2) or one fixed32 (wt 5)
        elif num in (5, 7):
            ints += [_signed(x) for x in _packed_varints(v, wt)]
        elif num == 8:
            name = v.decode()
        elif num == 9:
            raw = v
        elif num == 14 and v == 1:
            raise ValueError(f"tensor {name!r} keeps its data in an external file; merge it into the .onnx first")
    if dtype not in _DTYPES:
        raise ValueError(f"tensor {name!r}: ONNX data type {dtype} is not supported")
    dt = np.dtype(_DTYPES[dtype])
    if raw is not None:
        arr = np.frombuffer(raw, dtype=dt.newbyteorder("<")).astype(dt)
    elif floats:
        arr = np.frombuffer(b"".join(floats), dtype="<f4").astype(dt)
    else:
        arr = np.asarray(ints, dtype=dt)
    return Tensor(name, arr.reshape(dims) if dims else arr.reshape(()))


def _node(buf: bytes) -> Node:
    nd = Node("", [], [])
    for num, wt, v in _fields(buf):
        if num == 1:
            nd.inputs.append(v.decode())
        elif num == 2:
            nd.outputs.append(v.decode())
        elif num == 4:
            nd.op_type = v.decode()
        elif num == 5:
            aname, ival = "", None
            for n2, _, v2 in _fields(v):
                if n2 == 1:
                    aname = v2.decode()
                elif n2 == 3:
                    ival = _signed(v2)
            if ival is not None:
                nd.ints[aname] = ival
    return nd


def load(path) -> OnnxModel:
    buf = Path(path).read_bytes()
    meta: Dict[str, str] = {}
    inits: Dict[str, np.ndarray] = {}
    nodes: List[Node] = []
    for num, _, v in _fields(buf):
        if num == 14:                                   # metadata_props (ref OfflineModel.cs:31-72 reads this map)
            k = val = ""
            for n2, _, v2 in _fields(v):
                if n2 == 1:
                    k = v2.decode()
                elif n2 == 2:
                    val = v2.decode()
            meta[k] = val
        elif num == 7:                                  # graph
            for n2, _, v2 in _fields(v):
                if n2 == 5:
                    t = _tensor(v2)
                    inits[t.name] = t.array
                elif n2 == 1:
                    nodes.append(_node(v2))
    return OnnxModel(meta, inits, nodes)


# ---- which initialiser is which weight ---------------------------------------------------------------------------------------
def _dequant(m: OnnxModel, name: str) -> Optional[np.ndarray]:
    """fp32 view of initialiser `name`; for the onnxruntime quantiser's `<x>_quantized` also uses `<x>_scale` / `<x>_zero_point`."""
    a = m.initializers.get(name)
    if a is None:
        return None
    if a.dtype in (np.int8, np.uint8):
        base = name[:-len("_quantized")] if name.endswith("_quantized") else name
        scale, zp = m.initializers.get(base + "_scale"), m.initializers.get(base + "_zero_point")
        if scale is None:
            raise ValueError(f"quantised tensor {name!r} has no {base}_scale initialiser")
        z = np.zeros((), np.float32) if zp is None else zp.astype(np.float32)
        return ((a.astype(np.float32) - z) * scale.astype(np.float32)).astype(np.float32)
    return a.astype(np.float32)


def _linears(m: OnnxModel) -> List[Tuple[np.ndarray, Optional[np.ndarray], str]]:
    """Every Linear of the graph in node order as (weight [out,in], bias [out] or None, the node's input name)."""
    out = []
    producer_of = {o: nd for nd in m.nodes for o in nd.outputs}
    for i, nd in enumerate(m.nodes):
        if nd.op_type == "Gemm":
            w = _dequant(m, nd.inputs[1])
            if w is None:
                continue
            if not nd.ints.get("transB", 0):
                w = w.T
            b = _dequant(m, nd.inputs[2]) if len(nd.inputs) > 2 else None
            out.append((np.ascontiguousarray(w), b, nd.inputs[0]))
        elif nd.op_type in ("MatMul", "MatMulInteger"):
            w = _dequant(m, nd.inputs[1])
            if w is None or w.ndim != 2:
                continue
            w = np.ascontiguousarray(w.T)
            bias = None
            frontier = set(nd.outputs)
            for nx in m.nodes[i + 1:i + 8]:              # Add (after Cast / Mul in quantised graphs) with an initialiser of width `out`
                if frontier & set(nx.inputs):
                    if nx.op_type == "Add":
                        for x in nx.inputs:
                            c = _dequant(m, x)
                            if c is not None and c.size == w.shape[0]:
                                bias = c.reshape(-1)
                        break
                    frontier |= set(nx.outputs)
            src = nd.inputs[0]
            while src in producer_of and producer_of[src].op_type in ("DynamicQuantizeLinear", "Cast"):
                src = producer_of[src].inputs[0]
            out.append((w, bias, src))
    return out


def read_decoder(path) -> Tuple[Dict[str, np.ndarray], Dict[str, str]]:
    m = load(path)
    emb = conv = None
    for nd in m.nodes:
        if nd.op_type == "Gather" and emb is None:
            emb = _dequant(m, nd.inputs[0])
        elif nd.op_type == "Conv" and conv is None:
            conv = _dequant(m, nd.inputs[1])
    lin = _linears(m)
    if emb is None or conv is None or not lin:
        raise ValueError(f"{path}: not a stateless-decoder graph (Gather -> Conv -> Relu -> Gemm expected)")
    w, b, _ = lin[-1]
    if b is None:
        b = np.zeros(w.shape[0], np.float32)
    return ({"emb": np.ascontiguousarray(emb), "conv_w": np.ascontiguousarray(conv), "dec_proj_w": w, "dec_proj_b": b.reshape(-1)},
            m.metadata)


def read_joiner(path) -> Tuple[Dict[str, np.ndarray], Dict[str, str]]:
    m = load(path)
    lin = _linears(m)
    if not lin:
        raise ValueError(f"{path}: no Gemm / MatMul with a constant weight found")
    w, b, _ = lin[-1]                                    # output_linear: the one fed by Tanh
    if b is None:
        b = np.zeros(w.shape[0], np.float32)
    return {"out_w": w, "out_b": b.reshape(-1)}, m.metadata


def read_encoder_proj(path, joiner_dim: int) -> Dict[str, np.ndarray]:
    """encoder_proj, folded by upstream into encoder.onnx: the last Linear whose output width is joiner_dim."""
    m = load(path)
    for w, b, _ in reversed(_linears(m)):
        if w.shape[0] == joiner_dim:
            return {"enc_proj_w": w, "enc_proj_b": (b if b is not None else np.zeros(joiner_dim, np.float32)).reshape(-1)}
    raise ValueError(f"{path}: no Linear with {joiner_dim} outputs")


@dataclass
class B200Weights:
    """What k2b_load_weights takes, plus the metadata contract (ref OfflineModel.cs:31-46)."""
    weights: Dict[str, Optional[np.ndarray]]
    context_size: int
    vocab_size: int
    joiner_dim: int
    decoder_dim: int
    encoder_dim: int
    metadata: Dict[str, str]

    @classmethod
    def FromOnnx(cls, decoderFilePath, joinerFilePath, encoderFilePath=None) -> "B200Weights":
        dec, dmeta = read_decoder(decoderFilePath)
        joi, jmeta = read_joiner(joinerFilePath)
        meta = {**jmeta, **dmeta}
        V, D = dec["emb"].shape
        J = joi["out_w"].shape[1]
        ctx = int(dmeta.get("context_size", dec["conv_w"].shape[2]))          # ref OfflineModel.cs:34-35
        vocab = int(dmeta.get("vocab_size", V))                               # ref :37-38
        jd = int(jmeta.get("joiner_dim", J))                                  # ref :41-46
        if (vocab, jd) != (V, J) or joi["out_w"].shape[0] != V or dec["dec_proj_w"].shape != (J, D) or dec["conv_w"].shape != (D, 4, ctx):
            raise ValueError("decoder.onnx / joiner.onnx disagree on vocab_size / joiner_dim / decoder_dim with their metadata")
        w: Dict[str, Optional[np.ndarray]] = {**dec, **joi, "enc_proj_w": None, "enc_proj_b": None}
        E = 0
        if encoderFilePath is not None:
            ep = read_encoder_proj(encoderFilePath, J)
            w.update(ep)
            E = ep["enc_proj_w"].shape[1]
        return cls(w, ctx, V, J, D, E, meta)


# ---- a minimal writer, for fixtures (tests/golden/make_onnx_fixtures.py) ---------------------------------------------------------
def _enc_varint(x: int) -> bytes:
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _ld(num: int, payload: bytes) -> bytes:
    return _enc_varint(num << 3 | 2) + _enc_varint(len(payload)) + payload


def _vi(num: int, x: int) -> bytes:
    return _enc_varint(num << 3) + _enc_varint(x)


def enc_tensor(name: str, arr: np.ndarray, raw: bool = True) -> bytes:
    code = {v: k for k, v in _DTYPES.items()}[arr.dtype.type]
    out = b"".join(_vi(1, d) for d in arr.shape) + _vi(2, code) + _ld(8, name.encode())
    if raw or arr.dtype != np.float32:
        out += _ld(9, arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    else:
        out += _ld(4, arr.astype("<f4").tobytes())                             # packed float_data
    return out


def enc_node(op: str, inputs, outputs, ints: Optional[Dict[str, int]] = None) -> bytes:
    out = b"".join(_ld(1, s.encode()) for s in inputs) + b"".join(_ld(2, s.encode()) for s in outputs) + _ld(4, op.encode())
    for k, v in (ints or {}).items():
        out += _ld(5, _ld(1, k.encode()) + _vi(3, v) + _vi(20, 2))
    return out


def enc_model(nodes: List[bytes], tensors: List[bytes], metadata: Dict[str, str]) -> bytes:
    graph = b"".join(_ld(1, n) for n in nodes) + _ld(2, b"g") + b"".join(_ld(5, t) for t in tensors)
    out = _vi(1, 8) + _ld(2, b"k2b200-fixture") + _ld(7, graph)
    for k, v in metadata.items():
        out += _ld(14, _ld(1, k.encode()) + _ld(2, v.encode()))
    return out
