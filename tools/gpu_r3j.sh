#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "narrowing" 2>&1 | grep -v "^  \|Warning" | tail -30
RML_HOST_NARROW_TRACE=1 timeout 300 python tools/e2e_probe.py 2>&1 | tail -24
