#!/bin/bash
# round 2 (fourth session), GPU call 13: the mini-batch's H2D copy as many small pieces spread over the step (diagnostic)
mkdir -p gpurun_out
run() { echo -n "$1: "; env $1 RS_BENCH_E2E_PHASES=1 timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; grep "e2e phases" gpurun_out/last.err | sed 's/bookkeeping.*backward/backward/; s/allreduce.*end ->/end ->/'; python -c "
import json; d=json.load(open('gpurun_out/last.json')); print('      value %.2f ms/step  e2e %.2f ms/step' % (d['ms_per_step'], d['e2e']['ms_per_step']))"; }
{
run "RS_STAGE_CHUNKS=1"
run "RS_STAGE_CHUNKS=40"
run "RS_STAGE_CHUNKS=160 RS_STAGE_SPREAD_MS=9"
run "RS_STAGE_CHUNKS=8 RS_STAGE_SPREAD_MS=4"
run "RS_STAGE_CHUNKS=1"
run "RS_STAGE_CHUNKS=40"
} | tee gpurun_out/r02d_sweep13.log
