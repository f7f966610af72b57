#!/bin/bash
# Bench the current libgvs.so under several environment settings (one gpurun call):
#   tools/env_sweep.sh "GVS_STAGGER=1" "GVS_STAGGER=1 GVS_SELECT_CTAS_PER_SM=2" ...
# An empty string runs the default configuration.  Output under gpurun_out/sweep/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sweep
i=0
for e in "$@"; do
  i=$((i + 1))
  env $e timeout 200 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/sweep/bench_$i.json 2> gpurun_out/sweep/bench_$i.err
  python - "$i" "$e" <<'PY' | tee -a gpurun_out/sweep/summary.txt
import json, sys
i, e = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/sweep/bench_{i}.json").read().strip().splitlines()[-1])
    print(f"[{e}] ms/step {d['ms_per_step']:.3f} value {d['value']:.0f} e2e {d['e2e']['value']:.0f}")
except Exception as ex:
    print(f"[{e}] no bench line: {ex}")
PY
done
