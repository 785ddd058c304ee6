#!/usr/bin/env python
"""Generates tests/golden/net_scripts.json from the REFERENCE itself: Net<CpuDC>::to_script (Net.inl:521-530) of nets built with the
reference's own layer constructors, through tests/cpp/_build/libref_adapter_f32.so (rad_net_script).  Run in the build container."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob

lib = ctypes.CDLL(os.path.join(ROOT, "tests", "cpp", "_build", "libref_adapter_f32.so")); lib.rad_last_error.restype = ctypes.c_char_p
NETS = {
    "lenet": ob.lenet_specs(),
    "mlp": [ob.full((1, 1, 784), 100, ob.SIGMOID), ob.full((1, 1, 100), 10, ob.SIGMOID)],
    "strided_padded": [ob.conv((3, 32, 32), (3, 4), 8, ob.TANH, (1, 2), (2, 1)), ob.pool((8, 16, 33), (2, 3), ob.POOL_MIN),
                       ob.pool((8, 8, 11), (2, 2), ob.POOL_AVERAGE), ob.full((8, 4, 5), 7, ob.LINEAR)],
}
out = {}
for name, specs in NETS.items():
    arr = (ob.LayerSpec * len(specs))(*specs)
    buf = ctypes.create_string_buffer(1 << 16)
    assert lib.rad_net_script(arr, len(specs), buf, len(buf)) == 0, lib.rad_last_error()
    out[name] = {"specs": [[getattr(s, f) for f, _ in ob.LayerSpec._fields_] for s in specs], "script": buf.value.decode()}
path = os.path.join(ROOT, "tests", "golden", "net_scripts.json")
json.dump(out, open(path, "w"), indent=1)
print(path, {k: len(v["script"]) for k, v in out.items()})
