#!/bin/bash
echo "== mega small2"; timeout 900 python tools/mega_small2.py 2>&1 | grep -v "^$" | tail -14
