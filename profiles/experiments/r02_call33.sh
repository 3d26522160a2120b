set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
NC="--no-cpu-baseline --no-parity-check --steps 8 --warmup 3"
for t in main fn main fn; do
if [ $t = main ]; then unset CLPT_LIB; else export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so; fi
python bench.py $NC > gpurun_out/fn_$t.json 2> gpurun_out/fn_$t.err; python -c "import json;d=json.load(open('gpurun_out/fn_$t.json'));print('$t c3',d['value'],d['ms_per_step'])"
done
for t in main fn; do
if [ $t = main ]; then unset CLPT_LIB; else export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so; fi
python bench.py --config c2 $NC > gpurun_out/fn2_$t.json 2> gpurun_out/fn2_$t.err; python -c "import json;d=json.load(open('gpurun_out/fn2_$t.json'));print('$t c2',d['value'],d['ms_per_step'])"
done
export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_fn.so
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_fn.txt 2>&1; tail -3 gpurun_out/pytest_fn.txt
python bench.py --steps 3 > gpurun_out/fn_full.json 2> gpurun_out/fn_full.err;  python -c "import json;d=json.load(open('gpurun_out/fn_full.json'));print(d['value'],d['parity']['words_differ'])"
echo done
