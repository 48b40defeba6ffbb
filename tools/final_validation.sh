set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -c 600 gpurun_out/bench1.json
timeout 400 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gather_features_kernel -s 3 -c 1 -f -o gpurun_out/gather_r1final python tools/gather_bench.py --iters 3 > gpurun_out/g.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -s 12 -c 1 -f -o gpurun_out/gemm_car2_fwd_r1final python tools/gemm_bench.py quick > gpurun_out/gm.log 2>&1
timeout 100 python tools/gather_bench.py --profile A 2>&1 | grep "gather_features\"" | tail -1
timeout 100 python tools/gather_bench.py --profile B 2>&1 | grep "gather_features" | tail -2
