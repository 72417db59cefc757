#!/bin/bash
# Round 2, first GPU call (one GPU): the whole GPU tier (with the device-scalar / graph / xpay tests that never ran on
# hardware and the n = 2^28 reductions), bench.py, CG with host / device scalars / graph, ncu of the n = 10 / 12 kernels.
set -u
OUT=gpurun_out/r2a
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1

timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest -m gpu rc=$?" | tee "$OUT/summary.txt"
tail -5 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"

timeout 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "bench ref rc=$?" | tee -a "$OUT/summary.txt"

for mode in host fused device device3 device_fused graph; do
  for rep in 1 2; do
    timeout 300 libnomp_b200/build/cg_poisson 131072 8 61 1e-30 $mode 20 --nomp-backend cuda --nomp-device 0 --nomp-verbose 1 \
      | tail -1 | sed "s/^/{\"ranks\": 1, \"rep\": $rep, \"run\": /; s/$/}/" >> "$OUT/cg_scalars.jsonl"
  done
done
echo "cg lines: $(wc -l < "$OUT/cg_scalars.jsonl")" | tee -a "$OUT/summary.txt"

for n in 10 12; do
  E=$((n == 10 ? 131072 : 65536))
  timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax$n" \
      python tools/run_kernel_once.py ax $n $E 0 5 > /dev/null 2>&1
  ncu -i "$OUT/ax$n.ncu-rep" --page raw --csv > "$OUT/ax$n.raw.csv" 2> /dev/null
  ncu -i "$OUT/ax$n.ncu-rep" --page source --csv > "$OUT/ax$n.source.csv" 2> /dev/null
  rm -f "$OUT/ax$n.ncu-rep"
done
ls -la "$OUT" | tee -a "$OUT/summary.txt"
