#!/bin/bash
# round 2 (fourth session), GPU call 11: the prefetched mini-batch's H2D copy gated behind the previous step (RS_PREFETCH_COPY_GATE)
mkdir -p gpurun_out
run() { echo -n "$1: "; env $1 RS_BENCH_E2E_PHASES=1 timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; grep "e2e phases" gpurun_out/last.err | sed 's/bookkeeping.*backward/backward/; s/allreduce.*end ->/end ->/'; python -c "
import json; d=json.load(open('gpurun_out/last.json')); print('      value %.2f ms/step  e2e %.2f ms/step' % (d['ms_per_step'], d['e2e']['ms_per_step']))"; }
{
run "RS_PREFETCH_COPY_GATE=0"
run "RS_PREFETCH_COPY_GATE=1"
run "RS_PREFETCH_COPY_GATE=0"
run "RS_PREFETCH_COPY_GATE=1"
} | tee gpurun_out/r02d_sweep11.log
