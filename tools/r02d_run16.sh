#!/bin/bash
# round 2 (fourth session), GPU call 16: short first / last time chunks in the forward pass only (RS_TC_RAMP=2)
mkdir -p gpurun_out
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline'].get('families', {})
    g = lambda p: ([v for k, v in f.items() if k.startswith(p)] or [{'ms_per_step': float('nan')}])[0]['ms_per_step']
    print('   %.2f ms/step  e2e %.2f; fwd %.2f bwd %.2f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], g('lstm_stack_forward'), g('lstm_stack_backward')))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
for r in 0 2 3 0 2; do echo "RS_TC_RAMP=$r"; RS_TC_RAMP=$r bench; done
} | tee gpurun_out/r02d_sweep16.log
