#!/bin/bash
# usage: tools/ncu_kernel.sh <kernel-regex> <out-name> [ENV=VALUE ...]  -- one `ncu --set full` capture (2 launches) of a kernel inside a short bench run
k=$1; out=$2; shift 2
env "$@" python bench.py --steps 3 --frames 64 --no-cpu-baseline --no-sustained > /dev/null 2>&1 && \
env "$@" ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/$out python bench.py --steps 3 --frames 64 --no-cpu-baseline --no-sustained > gpurun_out/$out.log 2>&1
tail -1 gpurun_out/$out.log
