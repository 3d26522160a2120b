set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -q -s > gpurun_out/pytest_gpu_build_r02b.txt 2>&1; tail -8 gpurun_out/pytest_gpu_build_r02b.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spread or cooperative or readback or progressive" > gpurun_out/pytest_gpu_r02b.txt 2>&1; tail -5 gpurun_out/pytest_gpu_r02b.txt
CLPT_LIB=$PWD/clpathtracer_b200/libclpt_ctl.so python profiles/experiments/shard_kernel_times.py > gpurun_out/r02_shard_ctl.txt 2>&1
python profiles/experiments/shard_kernel_times.py > gpurun_out/r02_shard_regroup.txt 2>&1
CLPT_REGROUP=0 python profiles/experiments/shard_kernel_times.py > gpurun_out/r02_shard_g2_noregroup.txt 2>&1
CLPT_WARPS_PER_PIXEL=1 python profiles/experiments/shard_kernel_times.py > gpurun_out/r02_shard_g1.txt 2>&1
CLPT_WARPS_PER_PIXEL=4 python profiles/experiments/shard_kernel_times.py > gpurun_out/r02_shard_g4.txt 2>&1
python bench.py > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err
python bench.py --config c1 > gpurun_out/r02_bench_c1_b.json 2> gpurun_out/r02_bench_c1_b.err
python bench.py --config c5 > gpurun_out/r02_bench_c5_gpu.json 2> gpurun_out/r02_bench_c5_gpu.err; tail -3 gpurun_out/r02_bench_c5_gpu.err
python bench.py --config c5 --anim-builder host > gpurun_out/r02_bench_c5_host.json 2> gpurun_out/r02_bench_c5_host.err
python bench.py --config c5 --grid 707 > gpurun_out/r02_bench_c5_gpu_1m.json 2> gpurun_out/r02_bench_c5_gpu_1m.err; tail -3 gpurun_out/r02_bench_c5_gpu_1m.err
echo done
