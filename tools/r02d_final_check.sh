#!/bin/bash
# round 2 (fourth session), last 1-GPU call: the whole GPU suite, smoke() and the default bench line on the final tree
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -2 | tee gpurun_out/r02d_final_gpu_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02d_final_smoke.log
echo "== bench"; timeout 600 python bench.py 2>gpurun_out/r02d_final_bench.err | tail -1 > gpurun_out/r02d_final_bench_cfg2_n1.json
python - <<'PY' | tee gpurun_out/r02d_final.txt
import json
d = json.load(open('gpurun_out/r02d_final_bench_cfg2_n1.json'))
print('cfg2: value %.1f utt/s (%.2f ms/step)  e2e %.1f (%.2f ms/step)  with error rate %.2f ms  launches %d  frac %.4f  cpu_baseline %.2f utt/s on %d threads' % (
    d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['with_error_rate']['ms_per_step'], d['gpu_launches'], d['roofline']['frac'],
    d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
PY
