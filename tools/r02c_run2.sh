#!/bin/bash
# round 2 (third session), GPU call 2: register-FFT fbank kernel, leaner CTC lattice (REDUX row maximum, running pointers),
# hoist sweep, cfg-4 A/B of the 16-row validated exchange over two full cycles of its 8 batches
mkdir -p gpurun_out
echo "== tests (features, ctc, front door)"; timeout 900 python -m pytest tests/test_gpu_features.py tests/test_gpu_ctc.py tests/test_gpu_zz_frontdoor.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02c_tests_run2.log
echo "== tests with RS_FBANK_OCC=1"; RS_FBANK_OCC=1 timeout 900 python -m pytest tests/test_gpu_features.py -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/r02c_tests_run2.log
{
echo "RS_FBANK_FFT=0"; RS_FBANK_FFT=0 python tools/fbank_time.py
echo "RS_FBANK_FFT=1 RS_FBANK_OCC=2"; python tools/fbank_time.py
echo "RS_FBANK_FFT=1 RS_FBANK_OCC=1"; RS_FBANK_OCC=1 python tools/fbank_time.py
} 2>&1 | tee gpurun_out/r02c_fbank_time_run2.log
echo "== ctc"; timeout 300 python tests/gpu_diag.py ctc 2>&1 | grep "ms per" | tee gpurun_out/r02c_ctc_diag_run2.log
bench() { timeout 400 python bench.py --no-cpu-baseline "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline']['families']
    print('   value %.1f %s  %.2f ms/step  e2e %.1f ; fwd %.2f ctc %.2f bwd %.2f ms; with_error_rate %.2f ms/step' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'],
          [v for k, v in f.items() if k.startswith('lstm_stack_forward')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('ctc')][0]['ms_per_step'],
          [v for k, v in f.items() if k.startswith('lstm_stack_backward')][0]['ms_per_step'], d['with_error_rate']['ms_per_step']))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
for h in 20 16 12 8; do echo "RS_TC_HOIST=$h"; RS_TC_HOIST=$h bench --steps 10 --warmup 3; done
for x in 0 1 0 1; do echo "cfg4 RS_TS_XCHG16=$x"; RS_TS_XCHG16=$x bench --config cfg4 --steps 16 --warmup 8; done
} | tee gpurun_out/r02c_sweep2.log
