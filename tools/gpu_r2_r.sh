#!/bin/bash
timeout 1200 python -m pytest tests/test_reference_adapter.py -x -q -m gpu -k "per_sample" 2>&1 | tail -12
