mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pointwise.py -x -q -m gpu -k "flow" > gpurun_out/r2t_flow.log 2>&1; tail -3 gpurun_out/r2t_flow.log
timeout 300 python tools/time_quadrature.py > gpurun_out/r2t_quadtime.log 2>&1; cat gpurun_out/r2t_quadtime.log
for cfg in "4 256" "5 256" "4 768" "5 768"; do set -- $cfg
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2t_launches_p$1_$2.csv python bench.py --config scalability_3d --p $1 --elements $2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-quadrature > gpurun_out/r2t_bench_ncu_p$1_$2.log 2>&1
done
ls -la gpurun_out | tail -8
