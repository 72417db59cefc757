#!/bin/bash
# Start one process per GPU of this node for a program written against the libnomp API:
#   tools/run_ranks.sh <ranks> <program> [args...]
# Rank r gets NOMP_COMM_SIZE / NOMP_COMM_RANK / NOMP_COMM_ID_FILE (a fresh file under /dev/shm) and
# "--nomp-backend cuda --nomp-device r"; rank 0's output goes to stdout, the others' to <id file>.out.<r>.
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
N=$1; shift
ID="/dev/shm/nomp-run-$$-$(date +%s)"
export NOMP_INSTALL_DIR="${NOMP_INSTALL_DIR:-$ROOT/libnomp_b200}"
pids=()
for ((r = N - 1; r >= 0; r--)); do
  if [ "$r" -eq 0 ]; then
    NOMP_COMM_SIZE=$N NOMP_COMM_RANK=$r NOMP_COMM_ID_FILE=$ID "$@" --nomp-backend cuda --nomp-device $r &
  else
    NOMP_COMM_SIZE=$N NOMP_COMM_RANK=$r NOMP_COMM_ID_FILE=$ID "$@" --nomp-backend cuda --nomp-device $r > "$ID.out.$r" 2>&1 &
  fi
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=$?; done
if [ $rc -ne 0 ]; then for ((r = 1; r < N; r++)); do echo "--- rank $r"; tail -5 "$ID.out.$r"; done; fi
rm -f "$ID" "$ID".*
exit $rc
