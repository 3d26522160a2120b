set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final3.txt 2>&1; tail -3 gpurun_out/pytest_gpu_final3.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final3.txt 2>&1; tail -1 gpurun_out/smoke_final3.txt
python bench.py > gpurun_out/g_c3.json 2> gpurun_out/g_c3.err
for c in c1 c2 c4 c5; do python bench.py --config $c > gpurun_out/g_$c.json 2> gpurun_out/g_$c.err; done
python bench.py --config c5 --grid 707 --steps 20 > gpurun_out/g_c5_1m.json 2> gpurun_out/g_c5_1m.err
python bench.py --mode normal --no-jitter --spp 1 --depth 2 > gpurun_out/g_modeA_sah.json 2> gpurun_out/g_modeA_sah.err
python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 > gpurun_out/g_modeA_ref.json 2> gpurun_out/g_modeA_ref.err
for f in c3 c1 c2 c4 c5 c5_1m modeA_sah modeA_ref; do python -c "
import json;d=json.load(open('gpurun_out/g_$f.json'));print('$f',d.get('value'),d.get('ms_per_step'),(d.get('e2e') or {}).get('value'),(d.get('parity') or {}).get('words_differ'),d.get('p99_ms'),d.get('streamed_p50_ms'),(d.get('cpu_baseline') or {}).get('value'))"; done
NC="--steps 2 --warmup 3 --no-cpu-baseline --no-parity-check"
python bench.py $NC > gpurun_out/plain_c3.json 2> gpurun_out/plain_c3.err &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_g_c3 python bench.py $NC > gpurun_out/ncu_g_c3.log 2>&1
python bench.py $NC > gpurun_out/plain_c3.json 2> gpurun_out/plain_c3.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g_launches.csv python bench.py $NC > gpurun_out/ncu_launches.log 2>&1
echo done
