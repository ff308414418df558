#!/bin/bash
# round 2 (fourth session), GPU call 17: CUDA_DEVICE_MAX_CONNECTIONS (hardware queues; ~10 streams carry the schedule)
mkdir -p gpurun_out
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline'].get('families', {})
    g = lambda p: ([v for k, v in f.items() if k.startswith(p)] or [{'ms_per_step': float('nan')}])[0]['ms_per_step']
    print('   %.2f ms/step  e2e %.2f  with error rate %.2f; fwd %.2f bwd %.2f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['with_error_rate']['ms_per_step'], g('lstm_stack_forward'), g('lstm_stack_backward')))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
echo "default"; bench
for n in 4 16 32; do echo "CUDA_DEVICE_MAX_CONNECTIONS=$n"; CUDA_DEVICE_MAX_CONNECTIONS=$n bench; done
echo "default"; bench
echo "CUDA_DEVICE_MAX_CONNECTIONS=32"; CUDA_DEVICE_MAX_CONNECTIONS=32 bench
} | tee gpurun_out/r02d_sweep17.log
