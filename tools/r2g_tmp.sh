set -u
OUT=gpurun_out/r2i
mkdir -p "$OUT"
export NOMP_INSTALL_DIR="$PWD/libnomp_b200"
timeout 1200 python -m pytest tests/test_gs_gpu.py -m gpu -q --timeout 600 > "$OUT/pytest_gs.log" 2>&1; echo "pytest gs rc=$?" | tee "$OUT/summary.txt"
tail -4 "$OUT/pytest_gs.log" | tee -a "$OUT/summary.txt"
for rep in 1 2; do
for k in warp4 warp8 group; do
  unset NOMPK_GS_KERNEL NOMPK_GS_ROWS
  if [ $k = group ]; then export NOMPK_GS_KERNEL=group; fi
  if [ $k = warp8 ]; then export NOMPK_GS_ROWS=8; fi
  python tools/gs_bench.py 8 64 64 64 30 | tail -1 | sed "s/^/{\"kernel\": \"$k\", \"run\": /; s/$/}/" >> "$OUT/gs_bench.jsonl"
  python tools/gs_bench.py 8 64 64 8 30 | tail -1 | sed "s/^/{\"kernel\": \"$k\", \"run\": /; s/$/}/" >> "$OUT/gs_bench.jsonl"
  python tools/gs_bench.py 10 40 40 40 30 | tail -1 | sed "s/^/{\"kernel\": \"$k\", \"run\": /; s/$/}/" >> "$OUT/gs_bench.jsonl"
done
done
unset NOMPK_GS_KERNEL NOMPK_GS_ROWS
timeout 600 ncu --set full --clock-control none -f -k regex:gs_local_warp_kernel -s 2 -c 1 -o "$OUT/gs_warp" python tools/gs_bench.py 8 64 64 64 3 --no-warmup > /dev/null 2>&1
ncu -i "$OUT/gs_warp.ncu-rep" --page raw --csv > "$OUT/gs_warp.raw.csv" 2> /dev/null; rm -f "$OUT/gs_warp.ncu-rep"
ls -la "$OUT" | tee -a "$OUT/summary.txt"
