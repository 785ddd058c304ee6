#!/bin/bash
mkdir -p gpurun_out
echo "== small"; timeout 600 python tools/mega_small.py 2>&1 | tail -12
echo "== debug 37"; timeout 600 python tools/mega_debug.py 37 2>&1 | grep -v "^$" | tail -40
