#!/bin/bash
# round 2 (fourth session), GPU call 6: staging copy on the CPU or the DMA: which one slows the backward pass
mkdir -p gpurun_out
run() { echo -n "$1: "; env $1 RS_BENCH_E2E_PHASES=1 timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 >/dev/null | grep "e2e phases"; }
{
run "RS_DIAG_STAGE=nocpu"
run "RS_DIAG_STAGE=nodma"
run "RS_DIAG_STAGE="
run "RS_DIAG_STAGE=nocpu"
run "RS_DIAG_STAGE=nodma"
} | tee gpurun_out/r02d_e2e_phases3.txt
