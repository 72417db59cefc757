#!/bin/bash
# Round 2, second GPU call (one GPU): whole GPU tier again (new tests), bench, Ax prefetch experiments (DRAM bytes + timing).
set -u
OUT=gpurun_out/r2b
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest -m gpu rc=$?" | tee "$OUT/summary.txt"
tail -8 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
for shape in "10 32768" "12 16384"; do
  set -- $shape
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ax_kernel \
      --csv --log-file "$OUT/dram_n$1.csv" python tools/ax_dram_probe.py $1 $2 0,7,8,21,22,23,30,31,32,33,34,35,36 > "$OUT/dram_n$1.log" 2>&1
done
AX_SHAPES=10:131072,10:262144,12:65536 AX_VARIANTS=0,7,8,21,22,23,30,31,32,33,34,35,36 AX_ROUNDS=5 timeout 900 python tools/ax_sweep.py axrobust \
    > "$OUT/ax_interleaved.jsonl" 2> "$OUT/ax_interleaved.err"
echo "ax sweep rc=$?" | tee -a "$OUT/summary.txt"
ls -la "$OUT" | tee -a "$OUT/summary.txt"
