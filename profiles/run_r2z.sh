#!/bin/bash
# Round 2, run Z (1 GPU): edge-shape tests with the real kernels, then the whole GPU suite.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_native_gpu.py -x -q -m gpu -k "edge_shapes" 2>&1 | tail -25 > gpurun_out/pytest_r2z_edge.log; cat gpurun_out/pytest_r2z_edge.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_r2z.log; cat gpurun_out/pytest_r2z.log
