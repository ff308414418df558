#!/bin/bash
# round 2 (fourth session), GPU call 2: weight planes packed beside the input dense (RS_TC_PACK_SIDE), feature kernels of the
# next mini-batch enqueued behind the forward pass (RS_PREFETCH_GATE)
mkdir -p gpurun_out
echo "== model + step-protocol tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02d_tests_run2.log
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline'].get('families', {})
    g = lambda p: ([v for k, v in f.items() if k.startswith(p)] or [{'ms_per_step': float('nan')}])[0]['ms_per_step']
    print('   value %.1f %s  %.2f ms/step  e2e %.1f (%.2f ms/step); fwd %.2f ctc %.2f bwd %.2f ms' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'],
          g('lstm_stack_forward'), g('ctc'), g('lstm_stack_backward')))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
echo "cfg2 PACK_SIDE=0 GATE=0"; RS_TC_PACK_SIDE=0 RS_PREFETCH_GATE=0 bench
echo "cfg2 PACK_SIDE=1 GATE=0"; RS_PREFETCH_GATE=0 bench
echo "cfg2 PACK_SIDE=0 GATE=1"; RS_TC_PACK_SIDE=0 bench
echo "cfg2 PACK_SIDE=1 GATE=1"; bench
cp gpurun_out/last.json gpurun_out/r02d_bench_cfg2_run2.json
echo "cfg2 PACK_SIDE=0 GATE=0"; RS_TC_PACK_SIDE=0 RS_PREFETCH_GATE=0 bench
echo "cfg2 PACK_SIDE=1 GATE=1"; bench
} | tee gpurun_out/r02d_sweep2.log
