set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_gb3.so
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -q -x > gpurun_out/pytest_gb3.txt 2>&1; tail -5 gpurun_out/pytest_gb3.txt
timeout 300 python bench.py --config c5 > gpurun_out/gb3_c5.json 2> gpurun_out/gb3_c5.err; tail -3 gpurun_out/gb3_c5.err; python -c "import json;d=json.load(open('gpurun_out/gb3_c5.json'));print(d['p50_ms'],d['p99_ms'],d['breakdown_p50_ms'],d['parity']['words_differ'])"
timeout 300 python bench.py --config c5 --grid 707 --steps 20 > gpurun_out/gb3_c5_1m.json 2> gpurun_out/gb3_c5_1m.err; python -c "import json;d=json.load(open('gpurun_out/gb3_c5_1m.json'));print(d['p50_ms'],d['p99_ms'],d['breakdown_p50_ms'],d['parity']['words_differ'])"
timeout 300 python profiles/experiments/gpu_build_timing.py 707 2236 > gpurun_out/gb3_timing.txt 2>&1; tail -3 gpurun_out/gb3_timing.txt
echo done
