#!/bin/bash
# round 2, GPU call 11: (a) localise the illegal access seen on rank 1 of the 2-GPU bench (rank 1's data on one GPU);
# (b) two-chain forward recurrent kernel: parity (model suite incl. trained weights), A/B; (c) full suite
mkdir -p gpurun_out
echo "== (a) rank-1 data, launch blocking"; RANK=1 WORLD_SIZE=1 CUDA_LAUNCH_BLOCKING=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | grep -v "^$" | tail -12 | cut -c1-300 | tee gpurun_out/r02_rank1_blocking.log
echo "== (b) model + train tests (two chains default)"; timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_trained.py -x -q -s 2>&1 | grep -i "passed\|failed\|error\|shipped\|argmax\|Error\|assert" | tail -30 | tee gpurun_out/r02_model_tests_run11.log
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  err-rate step %.2f ms  launches/step %d  fwd %.2f bwd %.2f ms  rec fwd %s bwd %s frac %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step'], ['%.2f' % x for x in r['launch_ms']['fwd']], ['%.2f' % x for x in r['launch_ms']['bwd']], r['frac']))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_TS_CHAINS=1 2>&1 | tee -a gpurun_out/r02_sweep11.log
cp gpurun_out/last.json gpurun_out/r02_bench_cfg2_run11.json
run RS_TS_CHAINS=0 2>&1 | tee -a gpurun_out/r02_sweep11.log
run RS_TS_CHAINS=1 RS_TC_CHUNK_FWD=128 2>&1 | tee -a gpurun_out/r02_sweep11.log
run RS_TS_CHAINS=1 RS_TC_PHASES=0 2>&1 | tee -a gpurun_out/r02_sweep11.log
echo "== cfg5"; timeout 300 python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('   cfg5 %.0f clips/s  %.2f ms/step  e2e %.0f  p50 %.1f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['latency_ms']['p50']))"
echo "== (c) full gpu suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests_run11.log
