set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_l2.so
NC="--no-cpu-baseline --no-parity-check --steps 8 --warmup 3"
for v in none 0:1.0 0:0.6 0:0.3 1:0.2 2:0.15; do
if [ $v = none ]; then unset CLPT_L2_PERSIST; else export CLPT_L2_PERSIST=$v; fi
python bench.py --config c4 $NC > gpurun_out/l2_c4_$v.json 2> gpurun_out/l2_c4_$v.err; grep "L2 window" gpurun_out/l2_c4_$v.err; python -c "import json;d=json.load(open('gpurun_out/l2_c4_$v.json'));print('c4 $v',d['value'],d['ms_per_step'])"
done
for v in none 0:1.0 1:0.5; do
if [ $v = none ]; then unset CLPT_L2_PERSIST; else export CLPT_L2_PERSIST=$v; fi
python bench.py $NC > gpurun_out/l2_c3_$v.json 2> gpurun_out/l2_c3_$v.err; python -c "import json;d=json.load(open('gpurun_out/l2_c3_$v.json'));print('c3 $v',d['value'],d['ms_per_step'])"
done
echo done
