#!/bin/bash
# round 2 (fourth session), GPU call 14: per-step times of the end-to-end loop (is the mean above the median?)
mkdir -p gpurun_out
RS_BENCH_E2E_PHASES=1 timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json
grep "^e2e" gpurun_out/last.err | tee gpurun_out/r02d_e2e_per_step.txt
python -c "
import json; d=json.load(open('gpurun_out/last.json')); print('value %.2f ms/step  e2e %.2f ms/step' % (d['ms_per_step'], d['e2e']['ms_per_step']))" | tee -a gpurun_out/r02d_e2e_per_step.txt
