#!/bin/bash
# round 2, GPU call 3: MN-major weight-gradient GEMMs (no transposes) -- unit test, model parity, A/B against the transposing path
mkdir -p gpurun_out
echo "== tc tests"; timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -s 2>&1 | grep -v "^$" | tail -30 | tee gpurun_out/r02_tc_tests_run3.log
echo "== model tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -x -q 2>&1 | tail -8 | tee gpurun_out/r02_model_tests_run3.log
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  err-rate step %.2f ms  launches/step %d  fwd %.2f bwd %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step']))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_TC_TN=1 2>&1 | tee -a gpurun_out/r02_sweep3.log
cp gpurun_out/last.json gpurun_out/r02_bench_cfg2_tn.json
run RS_TC_TN=0 2>&1 | tee -a gpurun_out/r02_sweep3.log
run RS_TC_TN=1 RS_TC_SIDE_CTAS=52 2>&1 | tee -a gpurun_out/r02_sweep3.log
run RS_TC_TN=1 RS_TC_DX_CTAS=24 2>&1 | tee -a gpurun_out/r02_sweep3.log
echo "== bench cfg5"; timeout 600 python bench.py --config cfg5 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg5_run3.json 2> gpurun_out/r02_bench_cfg5_run3.err; tail -c 1500 gpurun_out/r02_bench_cfg5_run3.err; head -c 2500 gpurun_out/r02_bench_cfg5_run3.json
