mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_render_kernel_traversal.py tests/test_instancing.py -m gpu -x -q > gpurun_out/x12_tests.log 2>&1; echo "rc=$?" >> gpurun_out/x12_tests.log; tail -3 gpurun_out/x12_tests.log
EXP_NAME=x12 EXP_VARIANTS="D P0 D P0" EXP_CASES="c2 c3 c4 nw" bash tools/exp_variants.sh
