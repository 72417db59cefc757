#!/bin/bash
# Evidence for one round, run on the GPU box from the repo root:  gpurun -- 'bash tools/profile_round.sh r01'
# Writes gpurun_out/<tag>/: the bench lines (normal runs), the ncu launch list of the same bench command, and one
# `ncu --set full` capture per hot kernel.  tools/summarize_ncu.py turns them into profiles/<tag>_ncu_summary.md.
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"

python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
python bench.py --steps 20 --warmup 3 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"

# launch list of the bench command (numbers printed under ncu are not bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/bench_under_ncu.log" 2>&1

NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax8" python tools/run_kernel_once.py ax 8 262144 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax10" python tools/run_kernel_once.py ax 10 262144 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax12" python tools/run_kernel_once.py ax 12 65536 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax6" python tools/run_kernel_once.py ax 6 524288 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/axdot8" python tools/run_kernel_once.py axdot 8 262144 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/axdot10" python tools/run_kernel_once.py axdot 10 262144 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax8eo" python tools/run_kernel_once.py axeo 8 262144 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax10eo" python tools/run_kernel_once.py axeo 10 262144 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/axdot10eo" python tools/run_kernel_once.py axdoteo 10 262144 0 5 > /dev/null 2>&1
$NCU -k regex:reduce_kernel -s 3 -c 1 -o "$OUT/dot" python tools/run_kernel_once.py reduce 1 268435456 0 5 > /dev/null 2>&1
$NCU -k regex:map_vec -s 3 -c 1 -o "$OUT/add" python tools/run_kernel_once.py map 0 268435456 0 5 > /dev/null 2>&1
$NCU -k regex:gs_local -s 2 -c 1 -o "$OUT/gs" python tools/gs_bench.py 8 64 64 64 3 --no-warmup > /dev/null 2>&1
# racecheck of the four Ax shapes, launch-overhead sweep (BASELINE configs[4]), CG with host / device scalars / graph
for n in 6 8 10 12; do
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/run_kernel_once.py ax $n 301 0 2 > "$OUT/racecheck.ax$n.log" 2>&1
  echo "racecheck ax n=$n rc=$? $(grep -E 'RACECHECK SUMMARY' "$OUT/racecheck.ax$n.log" | tail -1)" >> "$OUT/racecheck.txt"
done
libnomp_b200/build/launch_overhead --nomp-backend cuda --nomp-device 0 --nomp-verbose 1 > "$OUT/launch_overhead.jsonl" 2> "$OUT/launch_overhead.err"
for mode in host fused device device3 device_fused graph; do
  for rep in 1 2; do
    timeout 300 libnomp_b200/build/cg_poisson 131072 8 61 1e-30 $mode 20 --nomp-backend cuda --nomp-device 0 --nomp-verbose 1 \
      | tail -1 | sed "s/^/{\"ranks\": 1, \"rep\": $rep, \"run\": /; s/$/}/" >> "$OUT/cg_scalars.jsonl"
  done
done
# the raw counter page of every capture as CSV (what tools/summarize_ncu.py reads); the reports themselves (23 MB
# each with sources) stay on the box: gpurun_out is capped at 64 MiB
for r in "$OUT"/*.ncu-rep; do ncu -i "$r" --page raw --csv > "${r%.ncu-rep}.raw.csv" 2> /dev/null; done
for r in ax10 ax12; do ncu -i "$OUT/$r.ncu-rep" --page source --csv 2> /dev/null | gzip > "$OUT/$r.source.csv.gz"; done
# DRAM bytes of every Ax variant kept for profiling (the prefetch-window study of round 2)
V=0,7,22,23,30,31,32,33,34,35,36,37,38,39,40,41,42,43,48,50,52,54,56
for shape in "10 32768" "12 16384" "6 131072" "8 65536"; do
  set -- $shape
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ax_kernel \
      --csv --log-file "$OUT/dram_n$1.csv" python tools/ax_dram_probe.py $1 $2 $V > /dev/null 2>&1
done
AX_SHAPES=10:262144,12:65536,6:524288,8:262144 AX_VARIANTS=$V AX_ROUNDS=5 python tools/ax_sweep.py axrobust > "$OUT/ax_interleaved.jsonl" 2> /dev/null
rm -f "$OUT"/*.ncu-rep
ls -la "$OUT"
