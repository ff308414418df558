#!/bin/bash
# round 2, GPU call 6: fused dropout hops, cached weight planes, warp-aggregated beam select -- whole GPU suite, then bench A/B
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -i "passed\|failed\|error\|beam search T=\|cfg-2 shape backward\|rel-L2\|kernel_\|input_w\|output_w" | tail -40 | tee gpurun_out/r02_gpu_tests_run6.log
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  err-rate step %.2f ms  launches/step %d  fwd %.2f bwd %.2f ms  rec fwd %s bwd %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step'], ['%.2f' % x for x in r['launch_ms']['fwd']], ['%.2f' % x for x in r['launch_ms']['bwd']]))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_X=default 2>&1 | tee -a gpurun_out/r02_sweep6.log
cp gpurun_out/last.json gpurun_out/r02_bench_cfg2_run6.json
run RS_TC_FUSE_DROPOUT=0 2>&1 | tee -a gpurun_out/r02_sweep6.log
run RS_BEAM_SHARE=1 2>&1 | tee -a gpurun_out/r02_sweep6.log
