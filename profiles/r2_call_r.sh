timeout 250 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --inflight 2 > gpurun_out/r2r_bench_inflight2.json 2> gpurun_out/r2r_bench_inflight2.err; tail -3 gpurun_out/r2r_bench_inflight2.err
timeout 250 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench_inflight1.json 2> gpurun_out/r2r_bench_inflight1.err
python - <<'PY'
import json
for f in ['inflight2','inflight1']:
    try:
        d=json.loads(open(f'gpurun_out/r2r_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), d['config']['launch'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
