#!/bin/bash
# round 2, GPU call 26: schedule parameters against the faster recurrent kernels
mkdir -p gpurun_out
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  launches/step %d  fwd %.2f bwd %.2f ms  rec fwd %s bwd %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step'], ['%.2f' % x for x in r['launch_ms']['fwd']], ['%.2f' % x for x in r['launch_ms']['bwd']]))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_TC_PHASES=0 2>&1 | tee -a gpurun_out/r02_sweep26.log
run RS_TC_CHUNK_FWD=128 2>&1 | tee -a gpurun_out/r02_sweep26.log
run RS_TC_CHUNK_FWD=64 2>&1 | tee -a gpurun_out/r02_sweep26.log
run RS_TC_CHUNK=96 2>&1 | tee -a gpurun_out/r02_sweep26.log
run RS_TC_CHUNK=160 2>&1 | tee -a gpurun_out/r02_sweep26.log
run RS_TC_WINDOW=3 2>&1 | tee -a gpurun_out/r02_sweep26.log
run RS_TC_PHASES=3 2>&1 | tee -a gpurun_out/r02_sweep26.log
