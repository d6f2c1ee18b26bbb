# fp32-accurate GEMMs on the tensor cores (split bf16 terms): parity + a bench line of the fp32 mode in each GEMM mode
timeout 500 python -m pytest tests/test_gpu_fp32_split.py -q -s > gpurun_out/r2w_fp32_split_tests.txt 2>&1; echo tests rc=$?; grep -E "passed|failed|Error|error|assert|GEMM|fp32-mode" gpurun_out/r2w_fp32_split_tests.txt | tail -25
for mode in x6 x3 simt; do
AVI_B200_FP32_GEMM=$mode timeout 300 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/r2w_bench_fp32_$mode.json 2> gpurun_out/r2w_bench_fp32_$mode.err; echo $mode rc=$?; tail -2 gpurun_out/r2w_bench_fp32_$mode.err
done
python - <<'PY'
import json
for f in ['x6','x3','simt']:
    try:
        d=json.loads(open(f'gpurun_out/r2w_bench_fp32_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), d['roofline']['kernel'], round(d['roofline']['achieved'],1), d['kernels_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
