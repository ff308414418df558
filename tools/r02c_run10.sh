#!/bin/bash
# round 2 (third session), GPU call 10: reduce-scatter of the two-chain backward kernel with 1 KB bulk copies (RS_TS_PUSH_BULK=1)
mkdir -p gpurun_out
echo "== tests (model, train) with RS_TS_PUSH_BULK=1"; RS_TS_PUSH_BULK=1 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02c_tests_run10.log
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline'].get('families', {})
    g = lambda p: ([v for k, v in f.items() if k.startswith(p)] or [{'ms_per_step': float('nan')}])[0]['ms_per_step']
    print('   value %.1f %s  %.2f ms/step  e2e %.1f; fwd %.2f ctc %.2f bwd %.2f ms; bwd launch ms %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'],
          g('lstm_stack_forward'), g('ctc'), g('lstm_stack_backward'), d['roofline']['launch_ms']['bwd']))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
echo "cfg2 RS_TS_PUSH_BULK=0"; bench
echo "cfg2 RS_TS_PUSH_BULK=1"; RS_TS_PUSH_BULK=1 bench
} | tee gpurun_out/r02c_sweep10.log
for b in 0 1; do echo "== xchg timeline RS_TS_PUSH_BULK=$b"; RS_TS_PUSH_BULK=$b timeout 300 python tests/gpu_diag.py xchg 2>&1 | grep -A40 "backward\|bwd" | head -70; done | tee gpurun_out/r02c_xchg_run10.txt
RS_TS_PUSH_BULK=1 timeout 300 python tests/gpu_diag.py trace 2>&1 | cut -c1-250 | tee gpurun_out/r02c_trace_run10.txt | head -8
