#!/bin/bash
# round 2, GPU call 13: beam search with two utterances per CTA (parity, time); backward GEMM SM split sweep
mkdir -p gpurun_out
echo "== ctc / beam tests"; timeout 900 python -m pytest tests/test_gpu_ctc.py tests/test_gpu_train.py -x -q -s 2>&1 | grep -i "passed\|failed\|beam search T=" | tail -6 | tee gpurun_out/r02_ctc_tests_run13.log
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  err-rate step %.2f ms  launches/step %d  fwd %.2f bwd %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step']))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_X=default 2>&1 | tee -a gpurun_out/r02_sweep13.log
cp gpurun_out/last.json gpurun_out/r02_bench_cfg2_run13.json
run RS_TC_DX_CTAS=32 2>&1 | tee -a gpurun_out/r02_sweep13.log
run RS_TC_DX_CTAS=32 RS_TC_SIDE_CTAS=36 2>&1 | tee -a gpurun_out/r02_sweep13.log
run RS_TC_DX_CTAS=48 RS_TC_SIDE_TPC=1 2>&1 | tee -a gpurun_out/r02_sweep13.log
run RS_TC_DX_CTAS=32 RS_TC_SIDE_TPC=1 2>&1 | tee -a gpurun_out/r02_sweep13.log
run RS_TC_DX_TPC=1 RS_TC_SIDE_TPC=1 2>&1 | tee -a gpurun_out/r02_sweep13.log
run RS_TC_DX_CTAS=24 RS_TC_SIDE_CTAS=40 2>&1 | tee -a gpurun_out/r02_sweep13.log
