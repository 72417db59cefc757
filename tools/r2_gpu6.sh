#!/bin/bash
# Round 2, sixth GPU call (one GPU): whole GPU tier on the new defaults, bench.
set -u
OUT=gpurun_out/r2f
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest -m gpu rc=$?" | tee "$OUT/summary.txt"
tail -8 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
ls -la "$OUT" | tee -a "$OUT/summary.txt"
