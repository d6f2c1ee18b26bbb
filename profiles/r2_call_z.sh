# packed fp32 (FFMA2) GELU in the conv0 / GEMM epilogues + lbs(pose2rot=False): parity, then the step
timeout 400 python -m pytest tests/test_gpu_kernels.py -q > gpurun_out/r2z_kernel_tests.txt 2>&1; echo kernel tests rc=$?; tail -3 gpurun_out/r2z_kernel_tests.txt
timeout 300 python -m pytest tests/test_gpu_path.py -q -k "lbs or wav2vec2 or predict_c1 or configs" > gpurun_out/r2z_path_tests.txt 2>&1; echo path tests rc=$?; tail -3 gpurun_out/r2z_path_tests.txt
for i in 1 2; do
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_$i.json 2> gpurun_out/r2z_bench_err.txt; echo bench rc=$?; tail -2 gpurun_out/r2z_bench_err.txt
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2z_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), d['kernels_ms_per_step'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
