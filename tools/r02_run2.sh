#!/bin/bash
# round 2, GPU call 2: whole GPU suite on the new step protocol, then the three bench configs and a short reference arm
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r02_gpu_tests_run2.log
echo "== bench cfg2"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg2_run2.json 2> gpurun_out/r02_bench_cfg2_run2.err; tail -c 1500 gpurun_out/r02_bench_cfg2_run2.err; head -c 3000 gpurun_out/r02_bench_cfg2_run2.json
echo; echo "== bench cfg5"; timeout 600 python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg5_run2.json 2> gpurun_out/r02_bench_cfg5_run2.err; tail -c 1500 gpurun_out/r02_bench_cfg5_run2.err; head -c 2500 gpurun_out/r02_bench_cfg5_run2.json
echo; echo "== bench cfg4"; timeout 600 python bench.py --config cfg4 --no-cpu-baseline > gpurun_out/r02_bench_cfg4_run2.json 2> gpurun_out/r02_bench_cfg4_run2.err; tail -c 1500 gpurun_out/r02_bench_cfg4_run2.err; head -c 2500 gpurun_out/r02_bench_cfg4_run2.json
echo; echo "== reference arm (short)"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -3 | tee gpurun_out/r02_bench_reference_short.json
