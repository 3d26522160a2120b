set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02a.txt 2>&1; tail -5 gpurun_out/pytest_gpu_r02a.txt
python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_r02.json > gpurun_out/ref_vs_cuda_r02.txt 2>&1
for t in coop3 coop4 coop6 coopmin4 coopmin16; do
  CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_r02_$t.json > gpurun_out/ref_vs_cuda_r02_$t.txt 2>&1
done
python tests/golden/ref_kernel_stats.py > gpurun_out/refstats_r02b.txt 2>&1
echo done
python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -3 gpurun_out/r02_bench_a.err
python bench.py --config c1 > gpurun_out/r02_bench_c1_a.json 2> gpurun_out/r02_bench_c1_a.err
python bench.py --config c1 --readback float4 --readback-sync --no-cpu-baseline > gpurun_out/r02_bench_c1_sync.json 2> gpurun_out/r02_bench_c1_sync.err
