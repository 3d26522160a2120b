set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python profiles/experiments/shard_kernel_times.py > gpurun_out/r02b_shard_regroup.txt 2>&1
CLPT_REGROUP=0 python profiles/experiments/shard_kernel_times.py > gpurun_out/r02b_shard_g2_noregroup.txt 2>&1
CLPT_WARPS_PER_PIXEL=1 python profiles/experiments/shard_kernel_times.py > gpurun_out/r02b_shard_g1.txt 2>&1
CLPT_LIB=$PWD/clpathtracer_b200/libclpt_ctl.so python profiles/experiments/shard_kernel_times.py > gpurun_out/r02b_shard_ctl.txt 2>&1
python profiles/experiments/readback_timing.py > gpurun_out/readback_timing.txt 2>&1
python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_r02b.json > gpurun_out/ref_vs_cuda_r02b.txt 2>&1
for t in live4 live16 live32; do
  CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_r02b_$t.json > gpurun_out/ref_vs_cuda_r02b_$t.txt 2>&1
done
python bench.py --config c5 --grid 707 --steps 30 > gpurun_out/r02_bench_c5_gpu_1m.json 2> gpurun_out/r02_bench_c5_gpu_1m.err; tail -3 gpurun_out/r02_bench_c5_gpu_1m.err
python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 --steps 5 --no-cpu-baseline --no-parity-check > gpurun_out/r02_modea_reftree.json 2> gpurun_out/r02_modea_reftree.err &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_modea_reftree python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 --steps 5 --no-cpu-baseline --no-parity-check > gpurun_out/ncu_r02_modea_reftree.log 2>&1
echo done
