set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02h.txt 2>&1; tail -4 gpurun_out/pytest_gpu_r02h.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.txt 2>&1; tail -1 gpurun_out/smoke_r02.txt
python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda_final.json > gpurun_out/ref_vs_cuda_final.txt 2>&1
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_bench_final.err | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
NC="--steps 2 --warmup 3 --no-cpu-baseline --no-parity-check"
python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 $NC > gpurun_out/plain_modea_ref.json 2> gpurun_out/plain_modea_ref.err &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_modea_ref python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 $NC > gpurun_out/ncu_r02_modea_ref.log 2>&1
echo done
