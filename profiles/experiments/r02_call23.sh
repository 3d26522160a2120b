set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
NC="--no-cpu-baseline --no-parity-check --steps 8 --warmup 3"
for t in xbase xpa xpl xbase xpl; do
CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so python bench.py $NC > gpurun_out/mv_$t.json 2> gpurun_out/mv_$t.err; python -c "import json;d=json.load(open('gpurun_out/mv_$t.json'));print('$t c3',d['value'],d['ms_per_step'])"
done
for t in xbase xpl; do
CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so python bench.py --config c4 $NC > gpurun_out/mv4_$t.json 2> gpurun_out/mv4_$t.err; python -c "import json;d=json.load(open('gpurun_out/mv4_$t.json'));print('$t c4',d['value'],d['ms_per_step'])"
CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so python bench.py --config c2 $NC > gpurun_out/mv2_$t.json 2> gpurun_out/mv2_$t.err; python -c "import json;d=json.load(open('gpurun_out/mv2_$t.json'));print('$t c2',d['value'],d['ms_per_step'])"
CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so python bench.py --mode normal --no-jitter --spp 1 --depth 2 --builder ref --tree-depth 15 $NC > gpurun_out/mvA_$t.json 2> gpurun_out/mvA_$t.err; python -c "import json;d=json.load(open('gpurun_out/mvA_$t.json'));print('$t modeA ref',d['value'],d['ms_per_step'])"
done
CLPT_LIB=$PWD/clpathtracer_b200/libclpt_xpl.so timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_xpl.txt 2>&1; tail -3 gpurun_out/pytest_xpl.txt
echo done
