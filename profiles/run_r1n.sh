#!/bin/bash
# Round-1 (session 8): L2 eviction-priority hints on the CTA-pair GEMM's TMA loads (FX_GEMM_L2HINT=<a><b>): DRAM bytes
# and time per launch of the four block-level GEMM shapes under ncu, then un-profiled timings and a short in-step run
# for the promising settings.
mkdir -p gpurun_out
for h in 00 02 12 10 22; do
  FX_GEMM_L2HINT=$h ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:gemm2 --csv --log-file gpurun_out/l2hint_$h.csv python tests/native/gemm_once.py > /dev/null 2>&1
done
python - <<'PY' | tee gpurun_out/l2hint_r1n.log
import csv, glob
for f in sorted(glob.glob("gpurun_out/l2hint_*.csv")):
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    iid, im, iv, iu = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = {}
    for r in rows:
        per.setdefault(r[iid], {})[r[im]] = (r[iv], r[iu])
    print(f)
    for k, m in per.items():
        print("  launch", k, {a: " ".join(b) for a, b in m.items()})
PY
for h in 00 02 12; do
  echo "== microbench FX_GEMM_L2HINT=$h" | tee -a gpurun_out/l2hint_r1n.log
  FX_GEMM_L2HINT=$h python tests/gpu_microbench.py gemm 2>&1 | tee -a gpurun_out/l2hint_r1n.log
done
for h in 00 02 12; do
  echo "== in-step FX_GEMM_L2HINT=$h" | tee -a gpurun_out/l2hint_r1n.log
  FX_GEMM_L2HINT=$h timeout 300 python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline 2>&1 | tail -1 | tee -a gpurun_out/l2hint_r1n.log
done
