#!/bin/bash
# round 2, GPU call 12: six-product input dense -- trained-weight parity, model suite, bench, full suite
mkdir -p gpurun_out
echo "== trained + model + train tests"; timeout 1200 python -m pytest tests/test_gpu_trained.py tests/test_gpu_model.py tests/test_gpu_train.py -x -q -s 2>&1 | grep -i "passed\|failed\|error\|shipped\|assert" | tail -12 | tee gpurun_out/r02_model_tests_run12.log
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  err-rate step %.2f ms  launches/step %d  fwd %.2f bwd %.2f ms  rec fwd %s bwd %s frac %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step'], ['%.2f' % x for x in r['launch_ms']['fwd']], ['%.2f' % x for x in r['launch_ms']['bwd']], r['frac']))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_X=default 2>&1 | tee -a gpurun_out/r02_sweep12.log
cp gpurun_out/last.json gpurun_out/r02_bench_cfg2_run12.json
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_gpu_tests_run12.log
