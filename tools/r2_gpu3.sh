#!/bin/bash
# Round 2, third GPU call (one GPU): the new Ax shared-memory layout (tests first), variant sweep with the prefetch eviction
# priorities, DRAM bytes per variant, ncu of the new n = 10 / 12 kernels, racecheck of the four shapes, bench.
set -u
OUT=gpurun_out/r2c
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_nomp_api_gpu.py -m gpu -q --timeout 600 -k "ax or Ax" > "$OUT/pytest_ax.log" 2>&1
echo "pytest ax rc=$?" | tee "$OUT/summary.txt"; tail -3 "$OUT/pytest_ax.log" | tee -a "$OUT/summary.txt"
V=0,7,22,23,30,31,32,33,34,35,36,37,38,39,40,41
for shape in "10 32768" "12 16384"; do
  set -- $shape
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ax_kernel \
      --csv --log-file "$OUT/dram_n$1.csv" python tools/ax_dram_probe.py $1 $2 $V > "$OUT/dram_n$1.log" 2>&1
done
AX_SHAPES=10:131072,12:65536,6:524288,8:262144 AX_VARIANTS=$V AX_ROUNDS=5 timeout 1200 python tools/ax_sweep.py axrobust \
    > "$OUT/ax_interleaved.jsonl" 2> "$OUT/ax_interleaved.err"
echo "ax sweep rc=$?" | tee -a "$OUT/summary.txt"
for n in 10 12; do
  E=$((n == 10 ? 131072 : 65536))
  for v in 0 23; do
    timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax${n}_v$v" \
        python tools/run_kernel_once.py ax $n $E $v 5 > /dev/null 2>&1
    ncu -i "$OUT/ax${n}_v$v.ncu-rep" --page raw --csv > "$OUT/ax${n}_v$v.raw.csv" 2> /dev/null
    ncu -i "$OUT/ax${n}_v$v.ncu-rep" --page source --csv 2> /dev/null | gzip > "$OUT/ax${n}_v$v.source.csv.gz"
    rm -f "$OUT/ax${n}_v$v.ncu-rep"
  done
done
for n in 6 8 10 12; do
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/run_kernel_once.py ax $n 301 0 2 > "$OUT/racecheck.ax$n.log" 2>&1
  echo "racecheck ax$n rc=$? $(grep -E 'RACECHECK SUMMARY' "$OUT/racecheck.ax$n.log" | tail -1)" | tee -a "$OUT/summary.txt"
done
timeout 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
ls -la "$OUT" | tee -a "$OUT/summary.txt"
