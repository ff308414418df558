#!/bin/bash
# round 2 (fourth session), GPU call 10: only the first chunk's rows of the input dense / the last chunk's rows of dtop in front
# of the first recurrent launch of a pass (RS_TC_SPLIT_INPUT, RS_TC_SPLIT_DTOP)
mkdir -p gpurun_out
echo "== model + step-protocol tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02d_tests_run10.log
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
for v in "0 0" "1 0" "0 1" "1 1" "0 0" "1 1"; do set -- $v; echo "RS_TC_SPLIT_INPUT=$1 RS_TC_SPLIT_DTOP=$2"; RS_TC_SPLIT_INPUT=$1 RS_TC_SPLIT_DTOP=$2 bench; done
} | tee gpurun_out/r02d_sweep10.log
