#!/bin/bash
# round 2 (fourth session), GPU call 5: which part of the end-to-end step slows its backward pass down
mkdir -p gpurun_out
run() { echo -n "$1: "; env $1 RS_BENCH_E2E_PHASES=1 timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 >/dev/null | grep "e2e phases"; }
{
run "RS_DIAG_E2E=a"
run "RS_DIAG_E2E=a RS_EARLY_READ=0"
run "RS_DIAG_E2E=c"
run "RS_DIAG_E2E=d"
run "RS_DIAG_E2E="
} | tee gpurun_out/r02d_e2e_phases2.txt
