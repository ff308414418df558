#!/bin/bash
# round 2, GPU call 14: slab colsum (bit-reproducible, faster), beam back to one utterance per CTA -- suite + bench + cfg4/cfg5 lines
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_gpu_tests_run14.log
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json')); r = d['roofline']; f = r['families']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  err-rate step %.2f ms  launches/step %d  fwd %.2f bwd %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d['gpu_launches'] / d['steps'], [v for k, v in f.items() if k.startswith('lstm_stack_f')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('lstm_stack_b')][0]['ms_per_step']))
except Exception as e:
    print('   FAILED', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
run RS_X=default 2>&1 | tee -a gpurun_out/r02_sweep14.log
cp gpurun_out/last.json gpurun_out/r02_bench_cfg2_run14.json
echo "== cfg5"; timeout 300 python bench.py --config cfg5 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg5_run14.json 2>/dev/null; python -c "
import json
d = json.load(open('gpurun_out/r02_bench_cfg5_run14.json')); print('   cfg5 %.0f clips/s  %.2f ms/step  e2e %.0f  p50 %.1f p95 %.1f ms  cpu %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['latency_ms']['p50'], d['e2e']['latency_ms']['p95'], d.get('cpu_baseline', {}).get('value')))"
echo "== cfg4"; timeout 400 python bench.py --config cfg4 > gpurun_out/r02_bench_cfg4_run14.json 2>/dev/null; python -c "
import json
d = json.load(open('gpurun_out/r02_bench_cfg4_run14.json')); print('   cfg4 %.1f utt/s  %.2f ms/step  e2e %.1f  err-rate step %.2f ms  cpu %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d.get('cpu_baseline', {}).get('value')))"
