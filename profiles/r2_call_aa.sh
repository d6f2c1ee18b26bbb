# programmatic dependent launch (griddepcontrol) of the GEMM / attention / LayerNorm kernels: parity with it on, then the step A/B
AVI_PDL=1 timeout 400 python -m pytest tests/test_gpu_kernels.py -q > gpurun_out/r2aa_kernel_tests.txt 2>&1; echo kernel tests rc=$?; tail -3 gpurun_out/r2aa_kernel_tests.txt
AVI_PDL=1 timeout 300 python -m pytest tests/test_gpu_path.py -q -k "graph or batched or wav2vec2 or predict_c1 or configs" > gpurun_out/r2aa_path_tests.txt 2>&1; echo path tests rc=$?; tail -3 gpurun_out/r2aa_path_tests.txt
for m in 1 0 1 0; do
AVI_PDL=$m timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2aa_bench_pdl${m}_$RANDOM.json 2> gpurun_out/r2aa_bench_err.txt; echo pdl $m rc=$?; tail -2 gpurun_out/r2aa_bench_err.txt
done
AVI_PDL=1 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/r2aa_bench_pdl1_inflight1.json 2>> gpurun_out/r2aa_bench_err.txt
AVI_PDL=0 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/r2aa_bench_pdl0_inflight1.json 2>> gpurun_out/r2aa_bench_err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2aa_bench_pdl*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['ms_per_step'],2), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e: print(f, 'ERR', e)
PY
