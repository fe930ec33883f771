mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest.log; tail -3 gpurun_out/r2i_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err; echo "bench c2 rc=$?"
for c in c1 c3 c4 c5 nw_final c3_instanced nw_final_instanced; do
  timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_$c.json 2> gpurun_out/r2i_bench_$c.err; echo "bench $c rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_ncu_launches.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY
