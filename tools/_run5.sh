mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_render_kernel_traversal.py tests/test_gpu_parity.py tests/test_random_scenes.py tests/test_instancing.py -m gpu -x -q > gpurun_out/x5_tests.log 2>&1; echo "rc=$?" >> gpurun_out/x5_tests.log; tail -4 gpurun_out/x5_tests.log
EXP_NAME=x5 EXP_VARIANTS="D B D" EXP_CASES="c2 c4 smoke" bash tools/exp_variants.sh
