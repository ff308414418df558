#!/bin/bash
# round 2, GPU call 27: forward tile fetched as two boxes
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_model.py -x -q -k "fault_injection or cfg1 or random_models or cfg2_shape_forward or cfg4" 2>&1 | tail -2
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  launches/step %d  fwd %.2f bwd %.2f ms  rec fwd %s bwd %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step'], ['%.2f' % x for x in r['launch_ms']['fwd']], ['%.2f' % x for x in r['launch_ms']['bwd']]))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_TS_SPLIT=1 2>&1 | tee -a gpurun_out/r02_sweep27.log
run RS_TS_SPLIT=0 2>&1 | tee -a gpurun_out/r02_sweep27.log
echo "== xchg diag split"; timeout 300 python tests/gpu_diag.py xchg 2>&1 | grep "rec ms\|cta0 chain 0" -A8 | grep -v bwd | head -12 | tee gpurun_out/r02_xchg_diag12.log
