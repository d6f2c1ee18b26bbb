for i in 1 2; do
AVI_GEMM_NO_MULTICAST=1 timeout 250 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_nomc$i.json 2> gpurun_out/r2l_bench_nomc$i.err
timeout 250 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_mc$i.json 2> gpurun_out/r2l_bench_mc$i.err
done
python - <<'PY'
import json
for f in ['nomc1','mc1','nomc2','mc2']:
    d=json.loads(open(f'gpurun_out/r2l_bench_{f}.json').read().strip().splitlines()[-1])
    print(f, round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['kernels_ms_per_step']['gemm_bf16_tc'], d['clocks'])
PY
