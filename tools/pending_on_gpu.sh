#!/bin/bash
# Everything that was written after the last GPU run and is waiting for its first one, in ONE gpurun call:
#   gpurun --timeout 1500 -- 'bash tools/pending_on_gpu.sh'           (one GPU, about ten minutes)
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/pending_on_gpu.sh 8' (adds the 8-rank CG comparison)
# Writes gpurun_out/pending/.  Order: correctness first (a failure there makes the timings meaningless).
#   1. the GPU tier, then the tests that skip until they have run on a device once (NOMP_RUN_PENDING=1);
#   2. examples/cg_poisson.c with host scalars against device scalars, E = 131072, N = 7;
#   3. interleaved Ax sweep of the variants that were only host-verified (21, 22, 23) against the defaults, n = 10 / 12;
#   4. ncu capture of the two-buffer n = 10 kernel that became the default without a profile.
set -u
RANKS=${1:-1}
OUT=gpurun_out/pending
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"

python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest -m gpu rc=$?" | tee "$OUT/summary.txt"
NOMP_RUN_PENDING=1 python -m pytest tests/test_device_scalars_gpu.py -m gpu -q > "$OUT/pytest_pending.log" 2>&1
echo "pending tests rc=$?" | tee -a "$OUT/summary.txt"

for mode in host fused device device3 device_fused; do
  for rep in 1 2 3; do   # interleaved: the boards are power-capped and drift by a few per cent between runs
    libnomp_b200/build/cg_poisson 131072 8 60 1e-30 $mode 20 --nomp-backend cuda --nomp-device 0 --nomp-verbose 1 \
      | tail -1 | sed "s/^/{\"ranks\": 1, \"rep\": $rep, \"run\": /; s/$/}/" >> "$OUT/cg_scalars.jsonl"
  done
done
if [ "$RANKS" -gt 1 ]; then
  for mode in host fused device device3 device_fused; do
    for rep in 1 2 3; do
      tools/run_ranks.sh "$RANKS" libnomp_b200/build/cg_poisson $((131072 / RANKS)) 8 60 1e-30 $mode 20 --nomp-verbose 1 \
        | tail -1 | sed "s/^/{\"ranks\": $RANKS, \"rep\": $rep, \"run\": /; s/$/}/" >> "$OUT/cg_scalars.jsonl"
    done
  done
fi
echo "cg host/device lines: $(wc -l < "$OUT/cg_scalars.jsonl")" | tee -a "$OUT/summary.txt"

AX_VARIANTS=0,7,8,21,22,23 AX_ROUNDS=9 python tools/ax_sweep.py axrobust > "$OUT/ax_interleaved.jsonl" 2> "$OUT/ax_interleaved.err"
echo "ax sweep rc=$?" | tee -a "$OUT/summary.txt"

ncu --set full --clock-control none --import-source on -f -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax10_twobuf" \
    python tools/run_kernel_once.py ax 10 131072 0 5 > /dev/null 2>&1
ncu -i "$OUT/ax10_twobuf.ncu-rep" --page raw --csv > "$OUT/ax10_twobuf.raw.csv" 2> /dev/null
rm -f "$OUT/ax10_twobuf.ncu-rep"
cat "$OUT/summary.txt"
