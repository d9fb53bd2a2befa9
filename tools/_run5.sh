mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; tail -8 gpurun_out/r2v_tests.log
