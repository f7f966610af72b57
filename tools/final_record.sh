#!/bin/bash
# The measurement record of a round (1 GPU), two gpurun calls (one ncu invocation per call):
#   tools/final_record.sh bench   parity suite, bench lines of every BASELINE.json workload, the reference arm, latency / temporal
#                                 probes, the ncu launch list
#   tools/final_record.sh full    a full-set ncu capture of one 64-frame step
# Output under gpurun_out/final/; tools/ncu_summary.py condenses full.ncu-rep into the table committed under profiles/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/final
mkdir -p $O
if [ "${1:-bench}" = full ]; then
# one step of our kernels at 64 frames (about 25 kernel launches per step: skip the three warm-up steps)
python bench.py --frames 64 --steps 1 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1 || { echo "bench failed without ncu"; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'pyr_fused|pyr_down|fast_cells|select_kernel|offsets_kernel|describe_fused|row_bucket|stereo_|grid_assign|tmp_|projb_' -s 75 -c 26 -f -o $O/full \
    python bench.py --frames 64 --steps 1 --warmup 3 --no-cpu-baseline --no-sustained > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
exit 0
fi
( time timeout 400 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log | head -1
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench_euroc.json 2> $O/bench_euroc.err; echo "bench euroc rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm rc=$?"
timeout 300 python bench.py --workload kitti --steps 10 --warmup 3 > $O/bench_kitti.json 2> $O/bench_kitti.err; echo "bench kitti rc=$?"
timeout 300 python bench.py --workload hires --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_hires.json 2> $O/bench_hires.err; echo "bench hires rc=$?"
timeout 300 python bench.py --workload mono --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_mono.json 2> $O/bench_mono.err; echo "bench mono rc=$?"
timeout 120 python tools/latency_probe.py > $O/latency.txt 2>&1; cat $O/latency.txt
timeout 120 python tools/temporal_probe.py > $O/temporal.txt 2>&1; tail -4 $O/temporal.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained > $O/ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], d.get("impl", "ours"), "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
              "roofline", d.get("roofline", {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "no line", e)
PY
