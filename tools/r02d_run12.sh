#!/bin/bash
# round 2 (fourth session), GPU call 12: cfg-5 with the next batch staged and copied under this batch's kernels
mkdir -p gpurun_out
timeout 600 python bench.py --config cfg5 --no-cpu-baseline 2>gpurun_out/r02d_bench_cfg5.err | tail -1 > gpurun_out/r02d_bench_cfg5_n1.json
python - <<'PY' | tee gpurun_out/r02d_cfg5.txt
import json
try:
    d = json.load(open('gpurun_out/r02d_bench_cfg5_n1.json'))
    e = d['e2e']
    print('cfg5: value %.0f clips/s (%.2f ms per batch); e2e pipelined %.0f clips/s (%.2f ms); one batch at a time %.0f clips/s (%.2f ms), p50 %.1f ms' % (
        d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['one_batch_at_a_time']['value'], e['one_batch_at_a_time']['ms_per_step'], e['latency_ms']['p50']))
except Exception as ex:
    print('failed', ex); print(open('gpurun_out/r02d_bench_cfg5.err').read()[-2000:])
PY
