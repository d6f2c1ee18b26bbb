# dynamic tile scheduling (cluster launch control) in the GEMM and conv0: kernel parity tests, graph / batching tests of the path,
# and the step A/B (default mask vs AVI_DYNAMIC_TILES=0)
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/r2y_kernel_tests.txt 2>&1; echo kernel tests rc=$?; tail -4 gpurun_out/r2y_kernel_tests.txt
timeout 300 python -m pytest tests/test_gpu_path.py -q -x -k "graph or batched or predict_fp32 or wav2vec2_bf16 or configs" > gpurun_out/r2y_path_tests.txt 2>&1; echo path tests rc=$?; tail -4 gpurun_out/r2y_path_tests.txt
for m in 15 0 15 0; do
AVI_DYNAMIC_TILES=$m timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench_dyn${m}_$RANDOM.json 2> gpurun_out/r2y_bench_err.txt; echo mask $m rc=$?; tail -2 gpurun_out/r2y_bench_err.txt
done
AVI_DYNAMIC_TILES=15 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/r2y_bench_dyn15_inflight1.json 2>> gpurun_out/r2y_bench_err.txt
AVI_DYNAMIC_TILES=0 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/r2y_bench_dyn0_inflight1.json 2>> gpurun_out/r2y_bench_err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2y_bench_dyn*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), d['kernels_ms_per_step']['gemm_bf16_tc'], d['kernels_ms_per_step']['conv0_gn_gelu'], d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
