set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02c.txt 2>&1; tail -5 gpurun_out/pytest_gpu_r02c.txt
python bench.py > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -3 gpurun_out/r02_bench_c.err
python bench.py --config c1 --steps 200 > gpurun_out/r02_bench_c1_c.json 2> gpurun_out/r02_bench_c1_c.err; tail -2 gpurun_out/r02_bench_c1_c.err | cut -c1-600
python bench.py --config c2 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
python bench.py --config c4 --steps 10 > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; tail -3 gpurun_out/r02_bench_c4_n1.err | cut -c1-600
python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_r02c.json > gpurun_out/ref_vs_cuda_r02c.txt 2>&1
for t in fat3 fat5 fat6; do
  CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_r02c_$t.json > gpurun_out/ref_vs_cuda_r02c_$t.txt 2>&1
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
echo done
