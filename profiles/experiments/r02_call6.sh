set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_build.py tests/test_gpu_parity.py -m gpu -q -k "build or fat or engine or spread" > gpurun_out/pytest_gpu_r02d.txt 2>&1; tail -4 gpurun_out/pytest_gpu_r02d.txt
B="--no-cpu-baseline --no-parity-check --steps 5"
for e in 1 2; do
  for c in c1 c2 c3 c4; do python bench.py --config $c --engine $e $B > gpurun_out/r02_eng${e}_$c.json 2> gpurun_out/r02_eng${e}_$c.err; done
  python bench.py --mode normal --no-jitter --spp 1 --depth 2 --engine $e $B > gpurun_out/r02_eng${e}_modea_sah.json 2>/dev/null
  python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 --engine $e $B > gpurun_out/r02_eng${e}_modea_ref.json 2>/dev/null
  python bench.py --spp 4 --depth 2 --grid 158 --engine $e $B > gpurun_out/r02_eng${e}_c5like.json 2>/dev/null
done
for m in 1 2; do
  CLPT_ROW_ORDER=$m python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 $B > gpurun_out/r02_order${m}_modea_ref.json 2>/dev/null
  CLPT_ROW_ORDER=$m python bench.py --mode normal --no-jitter --spp 1 --depth 2 $B > gpurun_out/r02_order${m}_modea_sah.json 2>/dev/null
  CLPT_ROW_ORDER=$m python bench.py --config c4 $B > gpurun_out/r02_order${m}_c4.json 2>/dev/null
  CLPT_ROW_ORDER=$m python bench.py --config c2 $B > gpurun_out/r02_order${m}_c2.json 2>/dev/null
done
python profiles/experiments/shard_kernel_times.py > gpurun_out/r02c_shard.txt 2>&1
CLPT_BUILD_TIMING_OPTS=--frames python profiles/experiments/gpu_build_timing.py 158 707 2236 > gpurun_out/gpu_build_timing.txt 2>&1
python bench.py --config c5 --steps 60 > gpurun_out/r02_bench_c5_gpu.json 2> gpurun_out/r02_bench_c5_gpu.err
python bench.py --config c5 --grid 707 --steps 60 > gpurun_out/r02_bench_c5_gpu_1m.json 2> gpurun_out/r02_bench_c5_gpu_1m.err; tail -3 gpurun_out/r02_bench_c5_gpu_1m.err
python profiles/experiments/gpu_build_timing.py 707 > gpurun_out/plain_build.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_build_launches.csv python profiles/experiments/gpu_build_timing.py 707 > gpurun_out/ncu_build.log 2>&1
echo done
