mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_pytest.log; tail -4 gpurun_out/r2h_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err; echo "bench rc=$?"; cat gpurun_out/r2h_bench_c2.json | cut -c1-600
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/r2h_k_render python tools/profile_frame.py 500 > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2h_ncu.log
