#!/bin/bash
# Evidence for one round, run on the GPU box from the repo root:  gpurun -- 'bash tools/profile_round.sh r01'
# Writes gpurun_out/<tag>/: the bench lines (normal runs), the ncu launch list of the same bench command, and one
# `ncu --set full` capture per hot kernel.  tools/summarize_ncu.py turns them into profiles/<tag>_ncu_summary.md.
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"

python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
python bench.py --steps 20 --warmup 3 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"

# launch list of the bench command (numbers printed under ncu are not bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/bench_under_ncu.log" 2>&1

NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax8" python tools/run_kernel_once.py ax 8 262144 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax10" python tools/run_kernel_once.py ax 10 131072 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax12" python tools/run_kernel_once.py ax 12 65536 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/ax6" python tools/run_kernel_once.py ax 6 524288 0 5 > /dev/null 2>&1
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/axdot8" python tools/run_kernel_once.py axdot 8 262144 0 5 > /dev/null 2>&1
$NCU -k regex:reduce_kernel -s 3 -c 1 -o "$OUT/dot" python tools/run_kernel_once.py reduce 1 268435456 0 5 > /dev/null 2>&1
$NCU -k regex:map_vec -s 3 -c 1 -o "$OUT/add" python tools/run_kernel_once.py map 0 268435456 0 5 > /dev/null 2>&1
# the raw counter page of every capture as CSV (what tools/summarize_ncu.py reads); the reports themselves (23 MB
# each with sources) stay on the box: gpurun_out is capped at 64 MiB
for r in "$OUT"/*.ncu-rep; do ncu -i "$r" --page raw --csv > "${r%.ncu-rep}.raw.csv" 2> /dev/null; done
for r in ax10 ax12; do ncu -i "$OUT/$r.ncu-rep" --page source --csv 2> /dev/null | gzip > "$OUT/$r.source.csv.gz"; done
rm -f "$OUT"/*.ncu-rep
ls -la "$OUT"
