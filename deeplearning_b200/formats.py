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
