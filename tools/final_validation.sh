set -x
timeout 300 python -m pytest tests -m gpu -q 2>&1 | grep -E "^FAILED|passed|failed" | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -c 300 gpurun_out/bench1.json
timeout 400 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 200 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
