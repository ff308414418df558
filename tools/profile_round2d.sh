#!/bin/bash
# One gpurun call at the end of round 2 (fourth session): the full GPU test suite, smoke(), the bench lines of the three
# configurations, the launch list of a bench run, ncu --set full of the two dominant recurrent kernels, the step trace.
# Usage (on the GPU box): bash tools/profile_round2d.sh <tag>
TAG=${1:-r02d}
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_gpu_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
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
echo "== trace"
timeout 300 python tests/gpu_diag.py trace > gpurun_out/${TAG}_trace.txt 2>&1; head -3 gpurun_out/${TAG}_trace.txt | cut -c1-200
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
