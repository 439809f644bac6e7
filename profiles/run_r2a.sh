#!/bin/bash
# Round 2, run A (1 GPU): parity suite incl. the config-2 shape tests, the bench line with the new parity /
# library_baseline / config-1 CPU legs, and a rasterisation sweep of the CTA-pair GEMM (DRAM bytes per launch under ncu).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2a.log 2>&1; tail -25 gpurun_out/pytest_gpu_r2a.log
timeout 900 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 6000 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
for cfg in "0 0" "12 6" "16 6" "24 6" "12 4" "24 0" "46 0" "6 0" "8 3" "1 6"; do
  set -- $cfg
  FX_GEMM_GROUP_M=$1 FX_GEMM_N_SPAN=$2 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:gemm2 --csv --log-file gpurun_out/raster_$1_$2.csv python tests/native/gemm_once.py > /dev/null 2>&1
done
python - <<'PY' | tee gpurun_out/raster_r2a.log
import csv, glob
for f in sorted(glob.glob("gpurun_out/raster_*.csv")):
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    if not rows: continue
    hdr, rows = rows[0], rows[1:]
    iid, im, iv = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = {}
    for r in rows:
        per.setdefault(r[iid], {})[r[im]] = r[iv]
    print(f)
    for (k, m), name in zip(per.items(), ("qkv", "o_resid", "ffn1", "ffn2")):
        rd = float(m["dram__bytes_read.sum"].replace(",", "")); wr = float(m["dram__bytes_write.sum"].replace(",", ""))
        print(f"  {name:8s} {float(m['gpu__time_duration.sum'].replace(',', ''))/1e3:9.1f} us  rd {rd/1e9:6.3f} GB  wr {wr/1e9:6.3f} GB  hit {m['lts__t_sector_hit_rate.pct']}")
PY
