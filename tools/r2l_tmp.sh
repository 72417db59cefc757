#!/bin/bash
set -u
OUT=gpurun_out/r2l
mkdir -p "$OUT"
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:ax_kernel -s 3 -c 1 -o "$OUT/axdot10p" python tools/run_kernel_once.py axdot 10 262144 63 5 > /dev/null 2>&1
ncu -i "$OUT/axdot10p.ncu-rep" --page raw --csv > "$OUT/axdot10p.raw.csv" 2> /dev/null
ncu -i "$OUT/axdot10p.ncu-rep" --page source --csv 2> /dev/null | gzip > "$OUT/axdot10p.source.csv.gz"
rm -f "$OUT"/*.ncu-rep
