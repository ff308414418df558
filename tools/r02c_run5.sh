#!/bin/bash
# round 2 (third session), GPU call 5: fbank kernel without the runtime cos/sin calls and with batched staging loads, dx GEMMs on
# up to (SMs - one recurrent launch) CTAs by default, cfg-5 with batch tiles of 32 (two-chain kernels) instead of 64
mkdir -p gpurun_out
echo "== tests (features)"; timeout 900 python -m pytest tests/test_gpu_features.py tests/test_gpu_zz_frontdoor.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02c_tests_run5.log
{ echo "RS_FBANK_OCC=2"; python tools/fbank_time.py; echo "RS_FBANK_OCC=1"; RS_FBANK_OCC=1 python tools/fbank_time.py; } 2>&1 | tee gpurun_out/r02c_fbank_time_run5.log
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline']['families']
    g = lambda p: ([v for k, v in f.items() if k.startswith(p)] or [{'ms_per_step': float('nan')}])[0]['ms_per_step']
    print('   value %.1f %s  %.2f ms/step  e2e %.1f (%s); fwd %.2f ctc %.2f bwd %.2f ms; with_error_rate %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'],
          d['e2e'].get('latency_ms', {}).get('p50'), g('lstm_stack_forward'), g('ctc'), g('lstm_stack_backward'), d.get('with_error_rate', {}).get('ms_per_step')))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
echo "cfg2 default"; bench
echo "cfg2 RS_TC_DX_CTAS=148"; RS_TC_DX_CTAS=148 bench
echo "cfg2 RS_TC_DX_CTAS=100 RS_TC_SIDE_CTAS=44"; RS_TC_DX_CTAS=100 RS_TC_SIDE_CTAS=44 bench
echo "cfg4 default"; bench --config cfg4 --steps 16 --warmup 8
echo "cfg5 tiles of 64"; bench --config cfg5
echo "cfg5 tiles of 32"; RS_TC_MAX_BATCH=32 bench --config cfg5
} | tee gpurun_out/r02c_sweep5.log
echo "== batch-tile tests with tiles of 32"; RS_TC_MAX_BATCH=32 timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "tiles or cfg5 or infer" 2>&1 | tail -3 | tee -a gpurun_out/r02c_tests_run5.log
