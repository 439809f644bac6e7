#!/bin/bash
# Round 2, run F (1 GPU): umT5 encoder with the tcgen05 attention kernel (operator tests, goldens, full depth), smoke(),
# the text-encoder bench leg.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "t5" > gpurun_out/pytest_gpu_r2f.log 2>&1; tail -30 gpurun_out/pytest_gpu_r2f.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2f.log 2>&1; tail -3 gpurun_out/smoke_r2f.log
timeout 600 python bench.py --workload t5 --steps 10 --warmup 3 > gpurun_out/t5_r2f.json 2> gpurun_out/t5_r2f.err; cat gpurun_out/t5_r2f.json; tail -3 gpurun_out/t5_r2f.err
timeout 300 python bench.py --workload t5 --steps 3 --warmup 2 --itemise > /dev/null 2>&1
