#!/bin/bash
# round 2, GPU call 16: self-validating exchange in the forward recurrent kernel (rec_ts_fwd3_kernel)
mkdir -p gpurun_out
echo "== model tests"; timeout 900 python -m pytest tests/test_gpu_model.py -x -q -s 2>&1 | grep -i "passed\|failed\|error\|near-tie\|utterances\|assert" | tail -12 | tee gpurun_out/r02_model_tests_run16.log
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  launches/step %d  fwd %.2f bwd %.2f ms  rec fwd %s bwd %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step'], ['%.2f' % x for x in r['launch_ms']['fwd']], ['%.2f' % x for x in r['launch_ms']['bwd']]))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_TS_XCHG=1 2>&1 | tee -a gpurun_out/r02_sweep16.log
run RS_TS_XCHG=0 2>&1 | tee -a gpurun_out/r02_sweep16.log
run RS_TS_XCHG=1 RS_TC_CHUNK_FWD=128 2>&1 | tee -a gpurun_out/r02_sweep16.log
