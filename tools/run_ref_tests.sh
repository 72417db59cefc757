#!/bin/bash
# Run the reference's own nomp-api test programs (built by `make -C oracle ref-tests` into oracle/_ref/tests)
# against this implementation, with the flag set of reference scripts/lnrun:120-130.
# The on-disk JIT cache is off unless the caller sets NOMP_JIT_CACHE=1 (with NOMP_JIT_CACHE_DIR): every kernel then
# goes through the transform bridge and NVRTC, as in the reference.
export NOMP_JIT_CACHE="${NOMP_JIT_CACHE:-0}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DIR="$ROOT/oracle/_ref/tests"
export NOMP_INSTALL_DIR="$ROOT/libnomp_b200"
cd "$DIR" || { echo "no reference tests built"; exit 2; }
for f in *.pyc.bin; do [ -f "$f" ] && cp -f "$f" "${f%.bin}"; done
fail=0
for t in nomp-api-*; do
  [ -x "$t" ] || continue
  timeout 300 ./"$t" --nomp-backend cuda --nomp-device 0 --nomp-platform 0 --nomp-install-dir "$NOMP_INSTALL_DIR" \
      --nomp-verbose "${NOMP_TEST_VERBOSE:-1}" --nomp-annotations-script sem
  rc=$?
  if [ $rc -eq 0 ]; then echo "$t: Passed"; else echo "$t: Failed (rc=$rc)"; fail=$((fail+1)); fi
done
echo "reference suite failures: $fail"
exit $fail
