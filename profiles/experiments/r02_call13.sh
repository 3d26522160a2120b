set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_lpr.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fat or engine or deterministic or counters or sah_trees or other_tree" > gpurun_out/pytest_gpu_lpr.txt 2>&1; tail -6 gpurun_out/pytest_gpu_lpr.txt | cut -c1-300
for k in 1 2 4; do
  CLPT_LANES_PER_RAY=$k python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_lpr$k.json > gpurun_out/ref_vs_cuda_lpr$k.txt 2>&1
done
unset CLPT_LIB
python -m pytest tests/test_c_host.py -m gpu -q -s -k gl_presentation > gpurun_out/pytest_c_host.txt 2>&1; tail -12 gpurun_out/pytest_c_host.txt | cut -c1-300
echo done
