set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_next.so
timeout 600 python -m pytest tests/test_gpu_build.py -m gpu -q > gpurun_out/pytest_gpu_build_next.txt 2>&1; tail -3 gpurun_out/pytest_gpu_build_next.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fat or engine or deterministic" > gpurun_out/pytest_gpu_next.txt 2>&1; tail -3 gpurun_out/pytest_gpu_next.txt
python profiles/experiments/gpu_build_timing.py 158 707 > gpurun_out/gpu_build_timing_next.txt 2>&1; tail -2 gpurun_out/gpu_build_timing_next.txt
python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_next.json > gpurun_out/ref_vs_cuda_next.txt 2>&1
CLPT_LIB=$PWD/clpathtracer_b200/libclpt_nopipe.so python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_nopipe.json > gpurun_out/ref_vs_cuda_nopipe.txt 2>&1
unset CLPT_LIB
python profiles/experiments/shard_c4_tiles.py > gpurun_out/shard_c4_tiles.txt 2>&1; tail -12 gpurun_out/shard_c4_tiles.txt
echo done
