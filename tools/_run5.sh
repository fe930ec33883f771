mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2i_bench_n2.json | head -c 1500; echo; tail -3 gpurun_out/r2i_bench_n2.err
timeout 120 python -m pytest tests/test_reduce.py tests/test_cpp_host.py -m gpu -x -q 2>&1 | tail -3
