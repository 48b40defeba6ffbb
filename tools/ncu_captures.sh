# `ncu --set full` captures of the kernels DESIGN.md quotes (one GPU; each replayed ~40x: run with `quick` sizes)
#   bash tools/ncu_captures.sh <tag>   -> gpurun_out/ncu_<kernel>_<tag>.ncu-rep + .txt summaries
tag=${1:-r2}
cap() { name=$1; shift; ncu --set full --clock-control none --import-source on -f -o gpurun_out/ncu_${name}_$tag "$@" > gpurun_out/ncu_${name}_$tag.log 2>&1; }
NAR_GEMM_BENCH_ONLY="bf16x3" cap gemm_bf16x3 -k regex:gemm_tf32_kernel -s 2 -c 1 python tools/gemm_bench.py quick
NAR_GEMM_BENCH_ONLY="B:MN +B_lo" cap gemm_3xtf32 -k regex:gemm_tf32_kernel -s 2 -c 1 python tools/gemm_bench.py quick
NAR_GEMM_BENCH_ONLY="dgrad 1x A:K  B:K" cap gemm_dgrad -k regex:gemm_tf32_kernel -s 2 -c 1 python tools/gemm_bench.py quick
cap gather_bulk -k regex:gather_features_kernel -s 3 -c 1 python tools/gather_bench.py --iters 2
cap segsum -k regex:car_segsum_kernel -s 1 -c 1 python tools/two_steps.py
cap combine -k regex:car_combine_kernel -s 1 -c 1 python tools/two_steps.py
for k in gemm_bf16x3 gemm_3xtf32 gemm_dgrad gather_bulk segsum combine; do
  python tools/summarize_ncu.py full gpurun_out/ncu_${k}_$tag.ncu-rep gpurun_out/ncu_${k}_$tag.txt $k || true
done
ls -la gpurun_out/ncu_*_$tag.* | head -30
