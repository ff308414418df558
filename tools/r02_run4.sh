#!/bin/bash
# round 2, GPU call 4: phase schedule (waves of L recurrent launches + GEMM bursts) -- parity, then A/B and chunk sweep
mkdir -p gpurun_out
echo "== model tests (phase schedule default)"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -x -q 2>&1 | tail -8 | tee gpurun_out/r02_model_tests_run4.log
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  err-rate step %.2f ms  launches/step %d  fwd %.2f bwd %.2f ms  rec fwd %s bwd %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step'], ['%.2f' % x for x in r['launch_ms']['fwd']], ['%.2f' % x for x in r['launch_ms']['bwd']]))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_TC_PHASES=1 2>&1 | tee -a gpurun_out/r02_sweep4.log
cp gpurun_out/last.json gpurun_out/r02_bench_cfg2_phases.json
run RS_TC_PHASES=0 2>&1 | tee -a gpurun_out/r02_sweep4.log
run RS_TC_PHASES=1 RS_TC_CHUNK=64 2>&1 | tee -a gpurun_out/r02_sweep4.log
run RS_TC_PHASES=1 RS_TC_CHUNK=96 2>&1 | tee -a gpurun_out/r02_sweep4.log
run RS_TC_PHASES=1 RS_TC_CHUNK=192 2>&1 | tee -a gpurun_out/r02_sweep4.log
run RS_TC_PHASES=1 RS_TC_SIDE_TPC=2 2>&1 | tee -a gpurun_out/r02_sweep4.log
run RS_TC_PHASES=1 RS_TC_DX_TPC=1 RS_TC_GX_TPC=1 2>&1 | tee -a gpurun_out/r02_sweep4.log
python tests/gpu_diag.py trace 2>&1 | tail -30 | tee gpurun_out/r02_trace_phases.txt
