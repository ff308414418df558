#!/bin/bash
# round 2 (fourth session), GPU call 18: what the CUDA events around every recurrent launch (the roofline's measurement) cost
# (run against the bench.py of that commit, where RS_BENCH_NO_TIMING=1 left the events out; since the next commit the
#  events have a timed region of their own and RS_BENCH_TIMED_EVENTS=1 puts them back into every region: the same A/B is
#  `RS_BENCH_TIMED_EVENTS=1 python bench.py` against `python bench.py`)
mkdir -p gpurun_out
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    print('   %.2f ms/step  e2e %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step']))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
for v in "" 1 "" 1; do echo "RS_BENCH_NO_TIMING=$v"; RS_BENCH_NO_TIMING=$v bench; done
} | tee gpurun_out/r02d_sweep18.log
