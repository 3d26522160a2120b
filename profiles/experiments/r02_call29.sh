set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
NC="--no-cpu-baseline --no-parity-check --steps 8 --warmup 3"
for t in main hd main hd; do
if [ $t = main ]; then unset CLPT_LIB; else export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so; fi
python bench.py $NC > gpurun_out/hd_$t.json 2> gpurun_out/hd_$t.err; python -c "import json;d=json.load(open('gpurun_out/hd_$t.json'));print('$t c3',d['value'],d['ms_per_step'],d['clocks'])"
done
for t in main hd; do
if [ $t = main ]; then unset CLPT_LIB; else export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_$t.so; fi
python bench.py --config c2 $NC > gpurun_out/hd2_$t.json 2> gpurun_out/hd2_$t.err; python -c "import json;d=json.load(open('gpurun_out/hd2_$t.json'));print('$t c2',d['value'],d['ms_per_step'])"
python bench.py --config c1 $NC > gpurun_out/hd1_$t.json 2> gpurun_out/hd1_$t.err; python -c "import json;d=json.load(open('gpurun_out/hd1_$t.json'));print('$t c1',d['value'],d['ms_per_step'])"
done
export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_hd.so
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_reference_kernel.py -m gpu -q -x > gpurun_out/pytest_hd.txt 2>&1; tail -3 gpurun_out/pytest_hd.txt
python bench.py --steps 3 > gpurun_out/hd_full.json 2> gpurun_out/hd_full.err;  python -c "import json;d=json.load(open('gpurun_out/hd_full.json'));print(d['value'],d['parity']['words_differ'],d['clocks'])"
echo done
