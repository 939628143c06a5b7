#!/bin/bash
# the long-read workload alone: parity tests of the long paths, then per-stage times through profiles/workloads.py
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_cli_gpu.py tests/test_sharded_gpu.py -k "long" -x -q --timeout=60 2>&1 | tail -3
