#!/bin/bash
# round 2 (fourth session), GPU call 20: cfg-4 bench line with the roofline's events in their own timed region
mkdir -p gpurun_out
timeout 100 python bench.py --config cfg4 --steps 8 --warmup 4 --no-cpu-baseline 2>gpurun_out/r02d_bench_cfg4.err | tail -1 > gpurun_out/r02d_bench_cfg4_n1_tworegions.json
python - <<'PY' | tee gpurun_out/r02d_cfg4_tworegions.txt
import json
try:
    d = json.load(open('gpurun_out/r02d_bench_cfg4_n1_tworegions.json'))
    print('cfg4: value %.1f utt/s (%.2f ms per batch)  e2e %.1f (%.2f ms)  under the events %.2f ms  frac %.4f' % (
        d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['instrumented_ms_per_step'], d['roofline']['frac']))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02d_bench_cfg4.err').read()[-1500:])
PY
