set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_gb.so
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -q -x > gpurun_out/pytest_gb.txt 2>&1; tail -5 gpurun_out/pytest_gb.txt
timeout 300 python profiles/experiments/gpu_build_timing.py 50 158 362 707 > gpurun_out/gb_timing.txt 2>&1; tail -5 gpurun_out/gb_timing.txt
CLPT_BUILD_NO_GRAPH=1 timeout 300 python profiles/experiments/gpu_build_timing.py 50 158 362 > gpurun_out/gb_timing_nograph.txt 2>&1; tail -4 gpurun_out/gb_timing_nograph.txt
timeout 300 python bench.py --config c5 > gpurun_out/gb_c5.json 2> gpurun_out/gb_c5.err; tail -3 gpurun_out/gb_c5.err; cat gpurun_out/gb_c5.json
CLPT_BUILD_NO_GRAPH=1 timeout 300 python bench.py --config c5 > gpurun_out/gb_c5_nograph.json 2> gpurun_out/gb_c5_nograph.err; cat gpurun_out/gb_c5_nograph.json
timeout 600 python -m pytest tests -m gpu -q -x --ignore=tests/test_gpu_build.py > gpurun_out/pytest_gb_rest.txt 2>&1; tail -5 gpurun_out/pytest_gb_rest.txt
echo done
