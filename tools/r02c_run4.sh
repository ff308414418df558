#!/bin/bash
# round 2 (third session), GPU call 4: full tests after the lattice fix (per-warp maxima of a one-warp block), more CTAs
# for the critical dx GEMMs, backward chunk length, ncu of the register-FFT fbank kernel and of the CTC lattice
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02c_gpu_tests_run4.log
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline']['families']
    print('   value %.1f %s  %.2f ms/step  e2e %.1f ; fwd %.2f ctc %.2f bwd %.2f ms; with_error_rate %.2f ms/step' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'],
          [v for k, v in f.items() if k.startswith('lstm_stack_forward')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('ctc')][0]['ms_per_step'],
          [v for k, v in f.items() if k.startswith('lstm_stack_backward')][0]['ms_per_step'], d['with_error_rate']['ms_per_step']))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
for c in 44 52 72 100; do echo "RS_TC_DX_CTAS=$c"; RS_TC_DX_CTAS=$c bench; done
for ch in 96 160; do echo "RS_TC_DX_CTAS=44 RS_TC_CHUNK=$ch RS_TC_CHUNK_FWD=96"; RS_TC_DX_CTAS=44 RS_TC_CHUNK=$ch RS_TC_CHUNK_FWD=96 bench; done
echo "RS_TC_DX_CTAS=44 RS_TC_SIDE_CTAS=60"; RS_TC_DX_CTAS=44 RS_TC_SIDE_CTAS=60 bench
} | tee gpurun_out/r02c_sweep4.log
RS_TC_DX_CTAS=44 timeout 300 python tests/gpu_diag.py trace > gpurun_out/r02c_trace_run4.txt 2>&1; cat gpurun_out/r02c_trace_run4.txt | cut -c1-250
echo "== ncu fbank"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbank_logmel2 -s 2 -c 1 -f -o gpurun_out/r02c_fbank_logmel2 python tools/fbank_time.py > gpurun_out/r02c_ncu_fbank.log 2>&1; tail -2 gpurun_out/r02c_ncu_fbank.log
echo "== ncu ctc"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_lattice1 -s 6 -c 1 -f -o gpurun_out/r02c_ctc_lattice1 python tests/gpu_diag.py ctc > gpurun_out/r02c_ncu_ctc.log 2>&1; tail -2 gpurun_out/r02c_ncu_ctc.log
ls -la gpurun_out/*.ncu-rep
