#!/bin/bash
# Round 2, run B (1 GPU): full parity suite with the round-2 engine (one host read per forward, token-mode timesteps,
# DenoiseLoop graphs, bf16x2 rmsnorm_rope, frame-wise GroupNorm partials, PDL on by default), bench line + itemised
# table, row-kernel microbenchmarks, the 50-step loop with and without CUDA graphs, ncu of the row kernels.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2b.log 2>&1; tail -40 gpurun_out/pytest_gpu_r2b.log
timeout 900 python bench.py --itemise > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 2500 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
FX_PDL=0 timeout 300 python bench.py --steps 5 --warmup 3 --quick > gpurun_out/bench_r2b_nopdl.json 2>&1; tail -c 600 gpurun_out/bench_r2b_nopdl.json
timeout 300 python tests/gpu_microbench.py rows > gpurun_out/micro_rows_r2b.log 2>&1; cat gpurun_out/micro_rows_r2b.log
timeout 600 python bench.py --workload loop50 > gpurun_out/loop50_r2b.json 2> gpurun_out/loop50_r2b.err; cat gpurun_out/loop50_r2b.json; tail -3 gpurun_out/loop50_r2b.err
timeout 600 python bench.py --workload loop50 --graph > gpurun_out/loop50_graph_r2b.json 2> gpurun_out/loop50_graph_r2b.err; cat gpurun_out/loop50_graph_r2b.json; tail -3 gpurun_out/loop50_graph_r2b.err
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:rmsnorm_rope' -c 2 \
    -o gpurun_out/prof_rows_r2b -f python tests/gpu_microbench.py rows > gpurun_out/prof_rows_r2b.log 2>&1
ncu -i gpurun_out/prof_rows_r2b.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread']
idx=[h.index(w) for w in want if w in h]
for r in rows[1:]: print([r[i] for i in idx])
" | tee gpurun_out/prof_rows_r2b.txt
