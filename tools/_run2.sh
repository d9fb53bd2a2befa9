mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pointwise.py -x -q -m gpu > gpurun_out/r2s_pointwise.log 2>&1; tail -5 gpurun_out/r2s_pointwise.log
timeout 300 python tools/time_quadrature.py > gpurun_out/r2s_quadtime.log 2>&1; cat gpurun_out/r2s_quadtime.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_brick --launch-skip 16 --launch-count 2 -o gpurun_out/r2s_brick -f python tools/time_quadrature.py 2:256 > gpurun_out/r2s_ncu.log 2>&1; tail -3 gpurun_out/r2s_ncu.log
timeout 300 python -m pytest tests/test_gpu.py -x -q -m gpu -k "quadrature" > gpurun_out/r2s_tests.log 2>&1; tail -5 gpurun_out/r2s_tests.log
