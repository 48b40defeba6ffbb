# ncu launch list of the bench command (cold-cache, serialised: compare SHARES) -> gpurun_out/launches_<tag>.csv / .md
tag=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$tag.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches_$tag.csv gpurun_out/launches_$tag.md || true
