#!/bin/bash
# Round 2, fifth GPU call (one GPU): one-element-per-group shapes against the production shapes, all with the local window.
set -u
OUT=gpurun_out/r2e
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -k "ax or Ax" > "$OUT/pytest_ax.log" 2>&1
echo "pytest ax rc=$?" | tee "$OUT/summary.txt"; tail -3 "$OUT/pytest_ax.log" | tee -a "$OUT/summary.txt"
V=0,7,30,31,33,35,39,42,43,44,45,46,47
AX_SHAPES=10:131072,12:65536,10:262144,6:524288,8:262144 AX_VARIANTS=$V AX_ROUNDS=5 timeout 1500 python tools/ax_sweep.py axrobust \
    > "$OUT/ax_interleaved.jsonl" 2> "$OUT/ax_interleaved.err"
echo "ax sweep rc=$?" | tee -a "$OUT/summary.txt"
for shape in "10 32768" "12 16384"; do
  set -- $shape
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__warps_active.avg.per_cycle_active,launch__registers_per_thread --clock-control none -k regex:ax_kernel \
      --csv --log-file "$OUT/dram_n$1.csv" python tools/ax_dram_probe.py $1 $2 $V > "$OUT/dram_n$1.log" 2>&1
done
# gather-scatter: where the time goes (L1 tag stage against DRAM)
timeout 600 ncu --set full --clock-control none -f -k regex:gs_local_kernel -s 2 -c 1 -o "$OUT/gs_local" python tools/gs_bench.py 8 64 64 64 > "$OUT/gs_bench.log" 2>&1
ncu -i "$OUT/gs_local.ncu-rep" --page raw --csv > "$OUT/gs_local.raw.csv" 2> /dev/null; rm -f "$OUT/gs_local.ncu-rep"
ls -la "$OUT" | tee -a "$OUT/summary.txt"
