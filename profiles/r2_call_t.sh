# round-2 evidence on the end-of-round code: default bench line, ncu launch list of the bench command (+ DRAM bytes per launch),
# one `ncu --set full` pass over the first 14 hot launches of a step (conv0, conv1-6, feature projection, pos-conv, layer 0's qkv /
# attention / out-proj / ffn1 / ffn2)
timeout 400 python bench.py > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; echo bench rc=$?; tail -2 gpurun_out/r2t_bench_n1.err
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
   --log-file gpurun_out/r2t_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --inflight 1 \
   > gpurun_out/r2t_ncu_bench.log 2>&1; echo launchlist rc=$?
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k 'regex:gemm_tc2_kernel|attn_tc_kernel|posconv_tc_kernel|conv0_tc_kernel' --launch-count 14 \
   -o gpurun_out/r2t_step_first14 -f python profiles/prof_step.py > gpurun_out/r2t_ncu_first14.log 2>&1; echo full rc=$?
ls -la gpurun_out/
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench_n1.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value']), d['e2e'], d['roofline'], d['clocks'], d.get('cpu_baseline'))
PY
