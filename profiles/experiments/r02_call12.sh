set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_c_host.py -m gpu -q -s > gpurun_out/pytest_c_host.txt 2>&1; tail -25 gpurun_out/pytest_c_host.txt | cut -c1-300
python -m pytest tests -m gpu -q --deselect tests/test_c_host.py::test_gl_presentation_hook > gpurun_out/pytest_gpu_r02g.txt 2>&1; tail -4 gpurun_out/pytest_gpu_r02g.txt
python bench.py --config c5 --steps 60 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; tail -2 gpurun_out/r02_bench_c5.err | cut -c1-300
echo done
