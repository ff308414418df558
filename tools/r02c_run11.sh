#!/bin/bash
# round 2 (third session), GPU call 11: in-kernel timeline of a backward launch INSIDE the schedule (two launches in flight + GEMMs)
# against one launch in flight + GEMMs
mkdir -p gpurun_out
{
timeout 300 python tests/gpu_diag.py xchg2
RS_TC_WINDOW_BWD=1 timeout 300 python tests/gpu_diag.py xchg2
RS_TC_DBG_LC=2,3 timeout 300 python tests/gpu_diag.py xchg2
} 2>&1 | tee gpurun_out/r02c_xchg2_run11.txt
