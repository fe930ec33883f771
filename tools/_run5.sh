mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_render_kernel_traversal.py -m gpu -x -q > gpurun_out/x11_tests.log 2>&1; echo "rc=$?" >> gpurun_out/x11_tests.log; tail -25 gpurun_out/x11_tests.log
