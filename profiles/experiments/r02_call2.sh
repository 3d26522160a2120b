set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -q --ignore=tests/test_gpu_build.py > gpurun_out/pytest_gpu_r02a.txt 2>&1; tail -5 gpurun_out/pytest_gpu_r02a.txt
timeout 600 python -m pytest tests/test_gpu_build.py -m gpu -q -s > gpurun_out/pytest_gpu_build_r02a.txt 2>&1; tail -5 gpurun_out/pytest_gpu_build_r02a.txt
( readelf -d /usr/local/nvidia/lib64/libEGL_nvidia.so.0 | grep -E "NEEDED|SONAME"; nm -D --defined-only /usr/local/nvidia/lib64/libEGL_nvidia.so.0 | head -40; ls /usr/local/nvidia/lib64 | grep -i -E "gldispatch|libEGL|libGL\.|libOpenGL|glvnd" ) > gpurun_out/egl_probe2.txt 2>&1
python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_r02.json > gpurun_out/ref_vs_cuda_r02.txt 2>&1
for t in coop3 coop4 coop6 coopmin4 coopmin16; do
  CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_r02_$t.json > gpurun_out/ref_vs_cuda_r02_$t.txt 2>&1
done
python tests/golden/ref_kernel_stats.py > gpurun_out/refstats_r02b.txt 2>&1
echo done
python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -3 gpurun_out/r02_bench_a.err
python bench.py --config c1 > gpurun_out/r02_bench_c1_a.json 2> gpurun_out/r02_bench_c1_a.err
python bench.py --config c1 --readback float4 --readback-sync --no-cpu-baseline > gpurun_out/r02_bench_c1_sync.json 2> gpurun_out/r02_bench_c1_sync.err
CLPT_WARPS_PER_PIXEL=1 python bench.py --no-cpu-baseline --no-parity-check > gpurun_out/r02_bench_g1.json 2> gpurun_out/r02_bench_g1.err
CLPT_WARPS_PER_PIXEL=4 python bench.py --no-cpu-baseline --no-parity-check > gpurun_out/r02_bench_g4.json 2> gpurun_out/r02_bench_g4.err
python profiles/experiments/shard_kernel_times.py > gpurun_out/r02_shard_g2.txt 2>&1
CLPT_WARPS_PER_PIXEL=1 python profiles/experiments/shard_kernel_times.py > gpurun_out/r02_shard_g1.txt 2>&1
CLPT_REGROUP=0 python bench.py --no-cpu-baseline --no-parity-check > gpurun_out/r02_bench_noregroup.json 2> gpurun_out/r02_bench_noregroup.err
CLPT_REGROUP=0 python profiles/experiments/shard_kernel_times.py > gpurun_out/r02_shard_g2_noregroup.txt 2>&1
