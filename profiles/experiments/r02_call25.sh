set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.txt 2>&1; tail -4 gpurun_out/pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1; tail -1 gpurun_out/smoke_final.txt
python bench.py > gpurun_out/f_c3.json 2> gpurun_out/f_c3.err; tail -2 gpurun_out/f_c3.err | cut -c1-300
for c in c1 c2 c4 c5; do python bench.py --config $c > gpurun_out/f_$c.json 2> gpurun_out/f_$c.err; done
python bench.py --config c5 --grid 707 --steps 20 > gpurun_out/f_c5_1m.json 2> gpurun_out/f_c5_1m.err
python bench.py --impl reference > gpurun_out/f_reference.json 2> gpurun_out/f_reference.err
python bench.py --mode normal --no-jitter --spp 1 --depth 2 > gpurun_out/f_modeA_sah.json 2> gpurun_out/f_modeA_sah.err
python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 > gpurun_out/f_modeA_ref.json 2> gpurun_out/f_modeA_ref.err
python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/f_ref_vs_cuda.json > gpurun_out/f_ref_vs_cuda.txt 2>&1
for f in c3 c1 c2 c4 c5 c5_1m modeA_sah modeA_ref; do python -c "
import json;d=json.load(open('gpurun_out/f_$f.json'));print('$f',d.get('value'),d.get('ms_per_step'),(d.get('e2e') or {}).get('value'),(d.get('parity') or {}).get('words_differ'))"; done
NC="--steps 2 --warmup 3 --no-cpu-baseline --no-parity-check"
python bench.py $NC > gpurun_out/plain_c3.json 2> gpurun_out/plain_c3.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py $NC > gpurun_out/ncu_launches.log 2>&1
prof() { # name, bench args
  name=$1; shift
  python bench.py "$@" $NC > gpurun_out/plain_$name.json 2> gpurun_out/plain_$name.err &&
  ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_f_$name python bench.py "$@" $NC > gpurun_out/ncu_f_$name.log 2>&1
}
prof c3
prof c4 --config c4
prof modeA_ref --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15
prof modeA_sah --mode normal --no-jitter --spp 1 --depth 2
prof c2 --config c2
prof c1 --config c1
ls -la gpurun_out/prof_f_*
echo done
