#!/bin/bash
# A/B harness for kernel variants on a GPU box (one gpurun call):  tools/ab_variants.sh <parity-variant> <variant> [<variant> ...]
# Every variant is a prebuilt variants/libgvs_<name>.so.  The first one also goes through the full GPU parity suite and a short
# differential fuzz; all of them get a short bench (stage times).  Output under gpurun_out/ab/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/ab
cp gtsam_vslam_b200/libgvs.so /tmp/libgvs_keep.so
first=$1
for v in "$@"; do
  cp variants/libgvs_$v.so gtsam_vslam_b200/libgvs.so
  if { [ "$v" = "$first" ] || [ "${PARITY_ALL:-0}" = 1 ]; } && [ "${SKIP_PARITY:-0}" != 1 ]; then
    ( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/ab/pytest_$v.log 2>&1; echo "pytest $v rc=$?" | tee -a gpurun_out/ab/summary.txt
    tail -3 gpurun_out/ab/pytest_$v.log
  fi
  if [ "$v" = "$first" ] && [ "${SKIP_PARITY:-0}" != 1 ]; then
    ( time timeout 200 python tools/fuzz_parity.py ${FUZZ_N:-400} 777 ) > gpurun_out/ab/fuzz_$v.log 2>&1; echo "fuzz $v rc=$?" | tee -a gpurun_out/ab/summary.txt
    tail -2 gpurun_out/ab/fuzz_$v.log
  fi
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab/bench_$v.json 2> gpurun_out/ab/bench_$v.err
  echo "bench $v rc=$?" | tee -a gpurun_out/ab/summary.txt
  python - "$v" <<'PY' | tee -a gpurun_out/ab/summary.txt
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab/bench_{v}.json").read().strip().splitlines()[-1])
    print(v, "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "stages", d.get("stage_ms"))
except Exception as e:
    print(v, "no bench line", e)
PY
done
cp /tmp/libgvs_keep.so gtsam_vslam_b200/libgvs.so
