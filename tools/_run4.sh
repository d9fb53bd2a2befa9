mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_tests.log 2>&1; tail -8 gpurun_out/r2u_tests.log
for cfg in "4 256" "5 256" "4 768" "5 768" "3 768" "2 768"; do set -- $cfg
timeout 600 python bench.py --config scalability_3d --p $1 --elements $2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-quadrature > gpurun_out/r2u_scal_p$1_$2.log 2>&1; tail -n 1 gpurun_out/r2u_scal_p$1_$2.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['config']['workload'][:40], round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['per_kernel'].items()}, round(d['roofline']['frac'],3))"
done
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2u_bench1.log 2>&1; tail -n 1 gpurun_out/r2u_bench1.log | cut -c1-1500
