#!/bin/bash
set -u
bash tools/profile_round.sh r02 > gpurun_out/profile_round.log 2>&1
echo "profile_round rc=$?"
ls gpurun_out/r02 | wc -l
