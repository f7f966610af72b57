#!/bin/bash
# usage: tools/ncu_kernel.sh <kernel-regex> <out-name> [skip] [count] [ENV=VALUE ...]  -- one `ncu --set full` capture of a kernel inside a short bench run
k=$1; out=$2; skip=${3:-8}; cnt=${4:-2}; shift 4 2>/dev/null || shift $#
env "$@" python bench.py --steps 3 --frames 64 --no-cpu-baseline --no-sustained > /dev/null 2>&1 && \
env "$@" ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -f -o gpurun_out/$out python bench.py --steps 3 --frames 64 --no-cpu-baseline --no-sustained > gpurun_out/$out.log 2>&1
tail -1 gpurun_out/$out.log | cut -c1-200
