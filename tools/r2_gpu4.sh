#!/bin/bash
# Round 2, fourth GPU call (one GPU): local prefetch windows (kPfMode 3 / 4) -- DRAM bytes and interleaved timings.
set -u
OUT=gpurun_out/r2d
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -k "ax or Ax" > "$OUT/pytest_ax.log" 2>&1
echo "pytest ax rc=$?" | tee "$OUT/summary.txt"; tail -3 "$OUT/pytest_ax.log" | tee -a "$OUT/summary.txt"
V=0,7,22,23,30,31,32,33,34,35,36,37,38,39,40,41
for shape in "10 32768" "12 16384" "6 131072"; do
  set -- $shape
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ax_kernel \
      --csv --log-file "$OUT/dram_n$1.csv" python tools/ax_dram_probe.py $1 $2 $V > "$OUT/dram_n$1.log" 2>&1
done
AX_SHAPES=10:131072,12:65536,6:524288,8:262144,10:262144 AX_VARIANTS=$V AX_ROUNDS=5 timeout 1500 python tools/ax_sweep.py axrobust \
    > "$OUT/ax_interleaved.jsonl" 2> "$OUT/ax_interleaved.err"
echo "ax sweep rc=$?" | tee -a "$OUT/summary.txt"
ls -la "$OUT" | tee -a "$OUT/summary.txt"
