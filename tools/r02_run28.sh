#!/bin/bash
# round 2, GPU call 28: beam search with the per-thread-maximum bound
mkdir -p gpurun_out
echo "== beam tests"; timeout 900 python -m pytest tests/test_gpu_ctc.py tests/test_gpu_trained.py -x -q -k "beam or trained" 2>&1 | tail -3 | tee gpurun_out/r02_beam_tests_run28.log
echo "== bench"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY' | tee gpurun_out/r02_sweep28.log
import json
d = json.load(open('gpurun_out/last.json'))
print('   value %.1f utt/s  %.2f ms/step  e2e %.1f ; with_error_rate %.1f utt/s %.2f ms/step' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['value'], d['with_error_rate']['ms_per_step']))
PY
echo "== beam kernel time"; ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ctc_beam_kernel -c 3 python -m pytest tests/test_gpu_ctc.py -q -k full_size 2>&1 | grep -i "ctc_beam_kernel\|gpu__time_duration" | head -8 | tee gpurun_out/r02_beam_time_run28.log
