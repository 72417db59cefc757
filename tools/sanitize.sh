#!/bin/bash
# compute-sanitizer passes over the hand-written kernels at small sizes (run on the GPU box from the repo root):
#   memcheck on every family, racecheck on the shared-memory kernels (Ax), synccheck on Ax.
# Prints one line per pass; exit code = number of passes with findings.
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
fail=0
run() {  # tool, tag, command...
  local tool=$1 tag=$2; shift 2
  timeout 900 compute-sanitizer --tool "$tool" --error-exitcode 9 "$@" > "$OUT/$tool.$tag.log" 2>&1
  local rc=$?
  local summary=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$OUT/$tool.$tag.log" | tail -1)
  echo "$tool $tag rc=$rc ${summary}"
  [ $rc -ne 0 ] && fail=$((fail+1))
}
for n in 6 8 10 12; do
  run memcheck ax$n python tools/run_kernel_once.py ax $n 301 0 2
  run racecheck ax$n python tools/run_kernel_once.py ax $n 301 0 2
done
run synccheck ax10 python tools/run_kernel_once.py ax 10 301 0 2
run memcheck axdot8 python tools/run_kernel_once.py axdot 8 301 0 2
run racecheck axdot10 python tools/run_kernel_once.py axdot 10 301 0 2
run memcheck dot python tools/run_kernel_once.py reduce 1 3000001 0 2
run memcheck add python tools/run_kernel_once.py map 0 3000001 0 2
run memcheck gs python tools/gs_bench.py 6 5 4 3 2 --no-warmup
run racecheck gs python tools/gs_bench.py 6 5 4 3 2 --no-warmup
run synccheck gs python tools/gs_bench.py 8 6 5 4 2 --no-warmup
NOMPK_GS_KERNEL=group run memcheck gs_group python tools/gs_bench.py 6 5 4 3 2 --no-warmup
# whole GPU test modules under memcheck: both gather-scatter kernels with emulated ranks, the two-level finish, the
# NVLink all-reduce on emulated ranks, the fused Ax / xpay kernels, device-resident scalars and graph replay
run memcheck t_gs python -m pytest tests/test_gs_gpu.py -m gpu -q -x -p no:cacheprovider
run synccheck t_gs python -m pytest tests/test_gs_gpu.py -m gpu -q -x -p no:cacheprovider -k "not setup"
run memcheck t_scalars python -m pytest tests/test_device_scalars_gpu.py -m gpu -q -x -p no:cacheprovider
run memcheck t_finish python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "two_level or emulated_ranks or cuda_graphs or fused_with_dot or mapped_host"
echo "sanitizer passes with findings: $fail"
exit $fail
