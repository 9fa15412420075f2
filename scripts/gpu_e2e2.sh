#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pinned_contexts or posterior_and_decision" 2>&1 | tail -8
