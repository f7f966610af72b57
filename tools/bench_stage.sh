#!/bin/bash
# usage: tools/bench_stage.sh <label> [ENV=VALUE ...] -- prints value / e2e / per-stage ms of a short bench run (A/B of tuning knobs)
label=$1; shift
env "$@" python bench.py --steps 10 --no-cpu-baseline --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$label', 'value %.0f ms %.3f e2e %.0f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']), d['stage_ms'])"
