#!/bin/bash
# round 2, GPU call 7: warp-level CTC lattice, beam search with bitonic sort of medium lists -- parity and time
mkdir -p gpurun_out
echo "== ctc / beam tests"; timeout 900 python -m pytest tests/test_gpu_ctc.py -x -q -s 2>&1 | grep -v "^$" | tail -15 | tee gpurun_out/r02_ctc_tests_run7.log
echo "== ctc old kernel"; RS_CTC_WARP=0 timeout 900 python -m pytest tests/test_gpu_ctc.py -x -q -s -k "cfg2 or shape or 998" 2>&1 | grep -v "^$" | tail -5
echo "== train tests"; timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_model.py -x -q 2>&1 | tail -4
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  err-rate step %.2f ms  launches/step %d  fwd %.2f ctc %.3f bwd %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('ctc')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step']))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_X=default 2>&1 | tee -a gpurun_out/r02_sweep7.log
cp gpurun_out/last.json gpurun_out/r02_bench_cfg2_run7.json
run RS_CTC_WARP=0 2>&1 | tee -a gpurun_out/r02_sweep7.log
