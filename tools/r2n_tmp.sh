#!/bin/bash
set -u
OUT=gpurun_out/r2n
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > "$OUT/pytest_gpu.log" 2>&1
echo "pytest -m gpu rc=$?" | tee "$OUT/summary.txt"; tail -3 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
bash tools/profile_round.sh r02 > "$OUT/profile_round.log" 2>&1
echo "profile_round rc=$?" | tee -a "$OUT/summary.txt"
AX_ROUNDS=5 AX_DOT_VARIANTS=61,64 AX_SHAPES=10:262144,12:65536,8:262144,6:524288 python tools/ax_sweep.py axdot > "$OUT/axdot_final.jsonl" 2> /dev/null
