#!/bin/bash
# round 2 (fourth session), GPU call 1: end_batch reads the accumulators behind the CTC kernel instead of draining the device
mkdir -p gpurun_out
echo "== step-protocol tests"; timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02d_tests_run1.log
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    print('   value %.1f %s  %.2f ms/step  e2e %.1f (%.2f ms/step)' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
echo "cfg2 RS_EARLY_READ=0"; RS_EARLY_READ=0 bench
echo "cfg2 RS_EARLY_READ=1"; bench
echo "cfg2 RS_EARLY_READ=0"; RS_EARLY_READ=0 bench
echo "cfg2 RS_EARLY_READ=1"; bench
} | tee gpurun_out/r02d_sweep1.log
cp gpurun_out/last.json gpurun_out/r02d_bench_cfg2_early.json
timeout 300 python tests/gpu_diag.py e2e 2>&1 | tail -5 | tee gpurun_out/r02d_e2e_diag.txt
