#!/bin/bash
# round 2 (third session), GPU call 3: layer 0's dx GEMM off the critical stream (RS_TC_DX0_SIDE), split of the spare SMs
mkdir -p gpurun_out
echo "== tests (model, train)"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02c_tests_run3.log
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
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
echo "RS_TC_DX0_SIDE=0"; RS_TC_DX0_SIDE=0 bench
echo "RS_TC_DX0_SIDE=1"; bench
for c in 24 32 44; do echo "RS_TC_DX0_SIDE=1 RS_TC_DX_CTAS=$c"; RS_TC_DX_CTAS=$c bench; done
for c in 36 52; do echo "RS_TC_DX0_SIDE=1 RS_TC_DX_CTAS=24 RS_TC_SIDE_CTAS=$c"; RS_TC_DX_CTAS=24 RS_TC_SIDE_CTAS=$c bench; done
echo "cfg4 RS_TC_DX0_SIDE=0"; RS_TC_DX0_SIDE=0 bench --config cfg4 --steps 16 --warmup 8
echo "cfg4 RS_TC_DX0_SIDE=1"; bench --config cfg4 --steps 16 --warmup 8
} | tee gpurun_out/r02c_sweep3.log
timeout 300 python tests/gpu_diag.py trace > gpurun_out/r02c_trace_run3.txt 2>&1; cat gpurun_out/r02c_trace_run3.txt | cut -c1-250
