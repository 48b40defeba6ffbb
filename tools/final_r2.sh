# last GPU call of round 2 (budget: ~5 min): representative parity subset on HEAD, one full ncu capture of the bf16x3 GEMM,
# the default bench line
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x --timeout 140 -k "tiny or gemm_bf16 or dedup_matches or device_resident or eval_ranking or aux_stream or sampler" > gpurun_out/pytest_r2z.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/pytest_r2z.log | cut -c1-300 | tail -8
NAR_GEMM_BENCH_ONLY="bf16x3" timeout 90 ncu --set full --clock-control none --import-source on -f -o gpurun_out/ncu_gemm_bf16x3_r2 -k regex:gemm_tf32_kernel -s 2 -c 1 python tools/gemm_bench.py quick > gpurun_out/ncu_gemm_bf16x3_r2.log 2>&1
python tools/summarize_ncu.py full gpurun_out/ncu_gemm_bf16x3_r2.ncu-rep gpurun_out/ncu_gemm_bf16x3_r2.txt "gemm_tf32_kernel<0,0,4,1,1>" 2>&1 | tail -2
rm -f gpurun_out/ncu_gemm_bf16x3_r2.ncu-rep
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err
tail -2 gpurun_out/bench_r2z.err | cut -c1-200
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r2z.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['host_enqueue_ms_per_step'], d['host_enqueue_ms_median_max'], d['roofline']['frac'], d['roofline_gather'].get('hbm_resident_form'))
PY
