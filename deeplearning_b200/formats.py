"""The reference's text format of a net's architecture: Net::to_script / Net::try_load_from_script
(NeuralNet/Net.inl:56-93, 521-541) - one line per layer,  "<CONV|PULL|FULL> <json>",  the json written by the layers' own to_script
(ALayer.inl:193-201 "Keep" / "Activation"; CLayer.inl:95-107 "InSize" "FilterSize" "FiltersCount" "Paddings" "Strides";
PLayer.inl:86-95 "InSize" "FilterSize" "PoolOperator"; NLayer.inl:74-81 "InSize" "OutSize"), vectors as "{a, b, c}" strings
(LinAlg3d.cpp Index3d::to_string), keys in nlohmann::json's alphabetical order, no spaces between tokens.

to_script writes exactly the text the reference writes for the same net; from_script accepts what the reference's reader accepts:
layer words in any case, missing "InSize" (the previous layer's output size is used), missing "Paddings", "Keep", "Activation", and
raises ValueError with the reference's messages for what it rejects."""
import json
import re
from typing import List, Optional, Sequence, Tuple

from ._capi import LayerSpec
from .net import ActivationFunctionId, CLayer, LayerTypeId, NLayer, PLayer, PoolTypeId, out_size

_TYPE_WORD = {LayerTypeId.CONVOLUTION: "CONV", LayerTypeId.FULL: "FULL", LayerTypeId.PULL: "PULL"}   # LayerTypeId.cpp:22-32


def _vec(values) -> str:
    return "{" + ", ".join(str(int(v)) for v in values) + "}"


def _parse_vec(text: str, n: int, what: str) -> Tuple[int, ...]:
    m = re.fullmatch(r"\s*\{\s*(-?\d+(?:\s*,\s*-?\d+)*)\s*\}\s*", text)
    vals = tuple(int(v) for v in re.split(r"\s*,\s*", m.group(1))) if m else ()
    if len(vals) != n:
        raise ValueError(f"Can't parse {what}: {text}")
    return vals


def layer_to_script(spec: LayerSpec, keep_rate: float = 1.0) -> str:
    t = LayerTypeId(spec.type)
    d = {"Keep": float(keep_rate)}
    if t == LayerTypeId.CONVOLUTION:
        d.update({"Activation": ActivationFunctionId(spec.activation).name, "InSize": _vec((spec.in_c, spec.in_h, spec.in_w)),
                  "FilterSize": _vec((spec.win_h, spec.win_w)), "FiltersCount": int(spec.out_c),
                  "Paddings": _vec((0, spec.pad_h, spec.pad_w)), "Strides": _vec((1, spec.stride_h, spec.stride_w))})
    elif t == LayerTypeId.PULL:
        d.update({"Activation": "UNKNOWN", "InSize": _vec((spec.in_c, spec.in_h, spec.in_w)), "FilterSize": _vec((spec.win_h, spec.win_w)),
                  "PoolOperator": PoolTypeId(spec.pool_type).name})
    elif t == LayerTypeId.FULL:
        d.update({"Activation": ActivationFunctionId(spec.activation).name, "InSize": _vec((1, 1, spec.in_c * spec.in_h * spec.in_w)),
                  "OutSize": _vec((1, 1, spec.out_c))})
    else:
        raise ValueError("Unknown layer type")
    return json.dumps(d, sort_keys=True, separators=(",", ":"))


def to_script(layers: Sequence[LayerSpec], keep_rates: Optional[Sequence[float]] = None) -> str:
    """Net::to_script (Net.inl:521-530)."""
    keep = list(keep_rates) if keep_rates is not None else [1.0] * len(layers)
    return "".join(_TYPE_WORD[LayerTypeId(s.type)] + " " + layer_to_script(s, k) + "\n" for s, k in zip(layers, keep))


def from_script(script: str) -> Tuple[List[LayerSpec], List[float]]:
    """Net::try_load_from_script (Net.inl:56-93): the layers and their keep rates."""
    layers: List[LayerSpec] = []
    keeps: List[float] = []
    prev_out = None
    for line in script.split("\n"):
        if not line.strip():
            continue
        m = re.match(r"\s*(\S+)\s*(.*)$", line, re.S)
        word = m.group(1).strip().upper().rstrip(";,")
        body = m.group(2)
        types = {v: k for k, v in _TYPE_WORD.items()}
        if word not in types:
            raise ValueError("Unknown layer type")
        t = types[word]
        try:
            js = json.loads(body)
        except json.JSONDecodeError as e:
            raise ValueError(f"Can't parse layer script: {line}") from e
        keep = float(js.get("Keep", 1.0))
        act_name = str(js.get("Activation", "UNKNOWN")).strip().upper()
        act = ActivationFunctionId[act_name] if act_name in ActivationFunctionId.__members__ else None
        in3 = _parse_vec(js["InSize"], 3, "Index3d") if "InSize" in js else prev_out
        if in3 is None:
            raise ValueError(f"Inconsistent input/output dimensions : {line}")
        if t == LayerTypeId.CONVOLUTION:
            if "FilterSize" not in js: raise ValueError("Can't parse filter window size of CLayer")
            if "FiltersCount" not in js: raise ValueError("Can't parse number of filters of CLayer")
            if act is None: raise ValueError("Unknown activation function ID")
            win = _parse_vec(js["FilterSize"], 2, "Index2d")
            pad = _parse_vec(js["Paddings"], 3, "Index3d") if "Paddings" in js else (0, 0, 0)
            if "Strides" not in js: raise ValueError("Can't parse strides of CLayer")   # (the reference falls back to {0, 0, 0} and divides by it)
            stride = _parse_vec(js["Strides"], 3, "Index3d")
            spec = CLayer(in3, win, int(js["FiltersCount"]), act, paddings=pad, strides=stride)
        elif t == LayerTypeId.PULL:
            if "FilterSize" not in js: raise ValueError("Can't parse window size of PLayer")
            if "PoolOperator" not in js: raise ValueError("Can't parse operator type of PLayer")
            op = str(js["PoolOperator"]).strip().upper()
            if op not in PoolTypeId.__members__: raise ValueError("Unknown pool operator")
            spec = PLayer(in3, _parse_vec(js["FilterSize"], 2, "Index2d"), PoolTypeId[op])
        else:
            if "OutSize" not in js: raise ValueError("Can't parse output dimensions of NLayer")
            if act is None: raise ValueError("Unknown activation function ID")
            o3 = _parse_vec(js["OutSize"], 3, "Index3d")
            spec = NLayer(in3[0] * in3[1] * in3[2], o3[0] * o3[1] * o3[2], act)
        if prev_out is not None:   # the sanity checks of Net.inl:81-89
            same = (prev_out[0] * prev_out[1] * prev_out[2] == spec.in_c * spec.in_h * spec.in_w) if t == LayerTypeId.FULL else \
                   (tuple(prev_out) == (spec.in_c, spec.in_h, spec.in_w))
            if not same:
                raise ValueError(f"Inconsistent input/output dimensions : {line}")
        layers.append(spec); keeps.append(keep)
        prev_out = out_size(spec)
    return layers, keeps


# ---------------------------------------------------------------------------------------------------------------------
# The reference's binary net files: Net::save_to_file / load_from_file (Net.inl:633-643) = msgpack (MsgPackUtils.h:27-70) of
#   Net          [ layers ]                                                        (MSGPACK_DEFINE(_layers), Net.h:111)
#   LayerHandle  [ layer type id, layer ]                                          (LayerHandle.inl:43-75)
#   ALayer       [ 1, keep rate, activation id ]                                   (ALayer.h:120-123)
#   CLayer       [ 1, ALayer, in size, weight tensor size, paddings, strides, biases, filters ]      (CLayer.inl:237-262)
#   PLayer       [ 1, ALayer, in size, pool window size, strides, pool operator id ]                 (PLayer.inl:36-60)
#   NLayer       [ 1, ALayer, biases, weights ]                                                      (NLayer.inl:96-120)
#   Index3d      [ x, y, z ];  Tensor [ layers, rows, columns, [data] ];  Matrix [ rows, columns, [data] ];  Vector [ [data] ]
# Real = float files carry 32-bit msgpack floats, Real = double files 64-bit ones.  msgpack-c itself is not in the reference tree
# (a NuGet dependency); what pins this reader and writer are the net files the reference ships for its own regression tests
# (DeepLearningTest/TestData/Regression/reference_net{,_single}.dat, read by NetTest.cpp:360-380): tests/test_formats.py parses
# them and requires save_net_bytes(load_net_bytes(file)) to reproduce each file byte for byte.
# ---------------------------------------------------------------------------------------------------------------------
def load_net_bytes(raw: bytes):
    """-> (layers, keep_rates, flat parameters in the library's order as a numpy array, "f32" | "f64")."""
    import msgpack
    import numpy as np
    obj = msgpack.unpackb(raw, raw=True, strict_map_key=False)
    if msgpack.packb(obj, use_single_float=True) == raw: real = "f32"
    elif msgpack.packb(obj, use_single_float=False) == raw: real = "f64"
    else: raise ValueError("Unexpected version of an object")   # (not a file this format writes)
    if not (isinstance(obj, list) and len(obj) == 1): raise ValueError("Unexpected version of an object")
    layers: List[LayerSpec] = []
    keeps: List[float] = []
    flat: List[float] = []
    for handle in obj[0]:
        type_id, body = handle
        if body[0] != 1 or body[1][0] != 1: raise ValueError("Unexpected version of an object")
        keep, act = float(body[1][1]), int(body[1][2])
        t = LayerTypeId(type_id)
        if t == LayerTypeId.CONVOLUTION:
            in3, w3, pad, stride, biases, filters = body[2], body[3], body[4], body[5], body[6], body[7]
            spec = CLayer(tuple(in3), (w3[1], w3[2]), len(filters), ActivationFunctionId(act), paddings=tuple(pad), strides=tuple(stride))
            flat.extend(biases[3])
            for f in filters: flat.extend(f[3])
        elif t == LayerTypeId.PULL:
            in3, w3, stride, op = body[2], body[3], body[4], body[5]
            if (w3[1], w3[2]) != (stride[1], stride[2]): raise ValueError("PLayer: strides other than the window are not supported")
            spec = PLayer(tuple(in3), (w3[1], w3[2]), PoolTypeId(op))
        elif t == LayerTypeId.FULL:
            biases, weights = body[2], body[3]
            spec = NLayer(int(weights[1]), int(weights[0]), ActivationFunctionId(act))
            flat.extend(biases[0]); flat.extend(weights[2])
        else:
            raise ValueError("Unknown layer type")
        layers.append(spec); keeps.append(keep)
    return layers, keeps, np.asarray(flat, np.float32 if real == "f32" else np.float64), real


def save_net_bytes(layers: Sequence[LayerSpec], keep_rates: Sequence[float], params, real: str = "f32") -> bytes:
    """Net::save_to_file's bytes for the given layers, keep rates and flat parameters (the library's order)."""
    import msgpack
    import numpy as np
    p = np.asarray(params, np.float32 if real == "f32" else np.float64)
    keep_dt = np.float32 if real == "f32" else np.float64
    off = 0
    handles = []
    for spec, keep in zip(layers, keep_rates):
        t = LayerTypeId(spec.type)
        base = [1, float(keep_dt(keep)), int(spec.activation)]
        if t == LayerTypeId.CONVOLUTION:
            oc, oh, ow = out_size(spec)
            nb, nf = oc * oh * ow, spec.in_c * spec.win_h * spec.win_w
            biases = [oc, oh, ow, p[off:off + nb].tolist()]; off += nb
            filters = []
            for _ in range(spec.out_c):
                filters.append([spec.in_c, spec.win_h, spec.win_w, p[off:off + nf].tolist()]); off += nf
            body = [1, base, [spec.in_c, spec.in_h, spec.in_w], [spec.in_c, spec.win_h, spec.win_w], [0, spec.pad_h, spec.pad_w],
                    [1, spec.stride_h, spec.stride_w], biases, filters]
        elif t == LayerTypeId.PULL:
            base[2] = 0   # a PLayer has no activation (ActivationFunctionId::UNKNOWN)
            body = [1, base, [spec.in_c, spec.in_h, spec.in_w], [1, spec.win_h, spec.win_w], [1, spec.win_h, spec.win_w], int(spec.pool_type)]
        elif t == LayerTypeId.FULL:
            n_in, n_out = spec.in_c * spec.in_h * spec.in_w, spec.out_c
            biases = [p[off:off + n_out].tolist()]; off += n_out
            weights = [n_out, n_in, p[off:off + n_out * n_in].tolist()]; off += n_out * n_in
            body = [1, base, biases, weights]
        else:
            raise ValueError("Unknown layer type")
        handles.append([int(spec.type), body])
    if off != p.size: raise ValueError("Invalid input")
    return msgpack.packb([handles], use_single_float=(real == "f32"))


def load_net_file(path):
    with open(path, "rb") as f:
        return load_net_bytes(f.read())


def save_net_file(path, layers, keep_rates, params, real: str = "f32") -> None:
    with open(path, "wb") as f:
        f.write(save_net_bytes(layers, keep_rates, params, real))
