#!/bin/bash
# round 2 (fourth session), GPU call 19: bench.py with the roofline's events in their own timed region
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/r02d_bench2.err | tail -1 > gpurun_out/r02d_bench_cfg2_n1_tworegions.json
python - <<'PY' | tee gpurun_out/r02d_tworegions.txt
import json
try:
    d = json.load(open('gpurun_out/r02d_bench_cfg2_n1_tworegions.json'))
    r = d['roofline']
    print('cfg2: value %.1f utt/s (%.2f ms/step)  e2e %.1f (%.2f ms/step)  with error rate %.2f ms  launches %d' % (
        d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['with_error_rate']['ms_per_step'], d['gpu_launches']))
    print('roofline: frac %.4f achieved %.1f avg_launch_ms %.3f share %.3f instrumented %.2f ms/step; %s' % (
        r['frac'], r['achieved'], r['avg_launch_ms'], r['share_of_step'], r['instrumented_ms_per_step'], r['measured_over']))
    print({k: round(v['ms_per_step'], 3) for k, v in r['families'].items() if isinstance(v, dict)})
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02d_bench2.err').read()[-2500:])
PY
