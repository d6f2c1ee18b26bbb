timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2s_gputests.txt 2>&1; tail -5 gpurun_out/r2s_gputests.txt
for c in 128 256; do
timeout 300 python bench.py --clips $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench_clips$c.json 2> gpurun_out/r2s_bench_clips$c.err; tail -2 gpurun_out/r2s_bench_clips$c.err
done
python - <<'PY'
import json
for c in (128,256):
    try:
        d=json.loads(open(f'gpurun_out/r2s_bench_clips{c}.json').read().strip().splitlines()[-1])
        print(c, round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']))
    except Exception as e: print(c,'ERR',e)
PY
