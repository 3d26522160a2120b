set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02e.txt 2>&1; tail -4 gpurun_out/pytest_gpu_r02e.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.txt 2>&1; tail -2 gpurun_out/smoke_r02.txt
NC="--steps 2 --warmup 3 --no-cpu-baseline --no-parity-check"
python bench.py $NC > gpurun_out/plain_c3.json 2> gpurun_out/plain_c3.err &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_c3 python bench.py $NC > gpurun_out/ncu_r02_c3.log 2>&1
python bench.py --config c4 $NC > gpurun_out/plain_c4.json 2> gpurun_out/plain_c4.err &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_c4 python bench.py --config c4 $NC > gpurun_out/ncu_r02_c4.log 2>&1
python bench.py --mode normal --no-jitter --spp 1 --depth 2 $NC > gpurun_out/plain_modea_sah.json 2> gpurun_out/plain_modea_sah.err &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_modea_sah python bench.py --mode normal --no-jitter --spp 1 --depth 2 $NC > gpurun_out/ncu_r02_modea_sah.log 2>&1
python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 $NC > gpurun_out/plain_modea_ref.json 2> gpurun_out/plain_modea_ref.err &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_modea_ref python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 $NC > gpurun_out/ncu_r02_modea_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/ncu_launches.log 2>&1
echo done
