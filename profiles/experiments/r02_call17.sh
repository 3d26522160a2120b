set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_k.so
timeout 600 python profiles/experiments/tail_probe.py 707 > gpurun_out/tail_probe.txt 2> gpurun_out/tail_probe.err; cat gpurun_out/tail_probe.txt; tail -3 gpurun_out/tail_probe.err
echo done
unset CLPT_LIB
NC="--no-cpu-baseline --no-parity-check --steps 6 --warmup 3"
for ci in 1 1.5 2 3; do
python bench.py --sah-ci $ci $NC > gpurun_out/sahci_$ci.json 2> gpurun_out/sahci_$ci.err; python -c "import json;d=json.load(open('gpurun_out/sahci_$ci.json'));print('ci',$ci,d['value'],d['ms_per_step'])"
done
echo done2
