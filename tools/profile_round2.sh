#!/bin/bash
# One gpurun call: launch list of a bench run + ncu --set full of the dominant kernels (round 2 names) + the step trace.
# Usage (on the GPU box): bash tools/profile_round2.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
for k in rec_ts_fwd2_kernel rec_ts_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 1 -f -o gpurun_out/${TAG}_$k \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|ctc_lattice_kernel|fbank_logmel_kernel|clip_adam_kernel|colsum_planes_kernel' -s 120 -c 30 -f -o gpurun_out/${TAG}_others \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_others.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_beam_kernel -c 1 -f -o gpurun_out/${TAG}_ctc_beam_kernel \
    python -m pytest tests/test_gpu_ctc.py -q -k full_size > gpurun_out/${TAG}_ctc_beam.log 2>&1
timeout 300 python tests/gpu_diag.py trace > gpurun_out/${TAG}_trace.txt 2>&1
ls -la gpurun_out | tail -20
