#!/bin/bash
# round 2 (fourth session), GPU call 9: the dgates planes' fill pattern written at the head of the backward pass (RS_TC_PREFILL)
mkdir -p gpurun_out
echo "== model + step-protocol tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02d_tests_run9.log
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline'].get('families', {})
    g = lambda p: ([v for k, v in f.items() if k.startswith(p)] or [{'ms_per_step': float('nan')}])[0]['ms_per_step']
    print('   %.2f ms/step  e2e %.2f; fwd %.2f bwd %.2f ms; launch ms fwd %s bwd %s' % (d['ms_per_step'], d['e2e']['ms_per_step'],
          g('lstm_stack_forward'), g('lstm_stack_backward'), ['%.2f' % x for x in d['roofline']['launch_ms']['fwd']], ['%.2f' % x for x in d['roofline']['launch_ms']['bwd']]))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
for v in 0 1 0 1; do echo "RS_TC_PREFILL=$v"; RS_TC_PREFILL=$v bench; done
echo "cfg4 RS_TC_PREFILL=0"; RS_TC_PREFILL=0 bench --config cfg4 --steps 8 --warmup 3
echo "cfg4 RS_TC_PREFILL=1"; RS_TC_PREFILL=1 bench --config cfg4 --steps 8 --warmup 3
} | tee gpurun_out/r02d_sweep9.log
