#!/bin/bash
# One gpurun call at the end of round 2 (third session): the full GPU test suite, the bench lines of the three configurations,
# the launch list of a bench run, ncu --set full of the kernels this session changed (register-FFT fbank, one-state-per-thread
# CTC lattice) and of the dominant recurrent kernels, the step traces.   Usage (on the GPU box): bash tools/profile_round2c.sh <tag>
TAG=${1:-r02c}
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_gpu_tests.log
echo "== bench cfg2"; timeout 600 python bench.py 2>gpurun_out/${TAG}_bench_cfg2.err | tail -1 > gpurun_out/${TAG}_bench_cfg2_n1.json; head -c 400 gpurun_out/${TAG}_bench_cfg2_n1.json; echo
echo "== bench cfg4"; timeout 600 python bench.py --config cfg4 --steps 16 --warmup 8 --no-cpu-baseline 2>gpurun_out/${TAG}_bench_cfg4.err | tail -1 > gpurun_out/${TAG}_bench_cfg4_n1.json; head -c 300 gpurun_out/${TAG}_bench_cfg4_n1.json; echo
echo "== bench cfg5"; timeout 600 python bench.py --config cfg5 --no-cpu-baseline 2>gpurun_out/${TAG}_bench_cfg5.err | tail -1 > gpurun_out/${TAG}_bench_cfg5_n1.json; head -c 300 gpurun_out/${TAG}_bench_cfg5_n1.json; echo
echo "== launches"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
for k in rec_ts_fwd3_kernel rec_ts_bwd4_kernel; do
  echo "== ncu $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 1 -f -o gpurun_out/${TAG}_$k \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_$k.log 2>&1
done
echo "== ncu fbank"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbank_logmel2 -s 2 -c 1 -f -o gpurun_out/${TAG}_fbank_logmel2 python tools/fbank_time.py > gpurun_out/${TAG}_ncu_fbank.log 2>&1
echo "== ncu ctc"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_lattice1 -s 18 -c 1 -f -o gpurun_out/${TAG}_ctc_lattice1 python tests/gpu_diag.py ctc > gpurun_out/${TAG}_ncu_ctc.log 2>&1
echo "== diags"
timeout 300 python tests/gpu_diag.py trace > gpurun_out/${TAG}_trace.txt 2>&1
RS_TRACE_CFG=4 timeout 300 python tests/gpu_diag.py trace > gpurun_out/${TAG}_trace_cfg4.txt 2>&1
timeout 300 python tools/fbank_time.py > gpurun_out/${TAG}_fbank_time.txt 2>&1
timeout 300 python tests/gpu_diag.py ctc 2>&1 | grep "ms per" > gpurun_out/${TAG}_ctc_time.txt
ls -la gpurun_out | grep ${TAG}
