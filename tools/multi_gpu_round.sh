#!/bin/bash
# Multi-GPU evidence of a round (gpurun --gpus 8 -- bash tools/multi_gpu_round.sh): the tests that need several GPUs, bench.py at N = 8 (and N = 2) launched the way
# the driver launches it, the CG example with host / device scalars / graph replay on 8 ranks.
set -u
OUT=gpurun_out/r2m
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "variants_agree" > "$OUT/pytest_variants.log" 2>&1
echo "pytest Ax variants rc=$?" | tee "$OUT/summary.txt"; tail -2 "$OUT/pytest_variants.log" | tee -a "$OUT/summary.txt"
timeout 1200 python -m pytest tests/test_system_gpu.py -m gpu -q --timeout 900 -k "two_gpus or across_gpus or graph_replay" > "$OUT/pytest_multi.log" 2>&1
echo "pytest multi-GPU rc=$?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/pytest_multi.log" | tee -a "$OUT/summary.txt"
for N in 8 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
  echo "bench N=$N rc=$?" | tee -a "$OUT/summary.txt"
done
for mode in host fused device3 device_fused graph; do
  timeout 300 tools/run_ranks.sh 8 libnomp_b200/build/cg_poisson 16384 8 61 1e-30 $mode 20 --nomp-verbose 1 \
    | tail -1 | sed "s/^/{\"ranks\": 8, \"run\": /; s/$/}/" >> "$OUT/cg_scalars.jsonl"
done
echo "cg lines: $(wc -l < "$OUT/cg_scalars.jsonl")" | tee -a "$OUT/summary.txt"
ls -la "$OUT" | tee -a "$OUT/summary.txt"
