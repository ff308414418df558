#!/bin/bash
# round 2 (third session), GPU call 9: input dropout in the input dense's epilogue, staging copy by four threads
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02c_gpu_tests_run9.log
bench() { timeout 400 python bench.py --no-cpu-baseline "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline'].get('families', {})
    g = lambda p: ([v for k, v in f.items() if k.startswith(p)] or [{'ms_per_step': float('nan')}])[0]['ms_per_step']
    print('   value %.1f %s  %.2f ms/step  e2e %.1f (%s); fwd %.2f ctc %.2f bwd %.2f ms; with_error_rate %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'],
          d['e2e'].get('latency_ms', {}).get('p50'), g('lstm_stack_forward'), g('ctc'), g('lstm_stack_backward'), d.get('with_error_rate', {}).get('ms_per_step')))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
echo "cfg2 default"; bench --steps 10 --warmup 3
echo "cfg2 RS_TC_FUSE_DROPOUT=0"; RS_TC_FUSE_DROPOUT=0 bench --steps 10 --warmup 3
echo "cfg4 default"; bench --config cfg4 --steps 16 --warmup 8
echo "cfg5 default"; bench --config cfg5 --steps 10 --warmup 3
} | tee gpurun_out/r02c_sweep9.log
