set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -q -x > gpurun_out/pytest_gb2.txt 2>&1; tail -5 gpurun_out/pytest_gb2.txt
NC="--no-cpu-baseline --no-parity-check --steps 6 --warmup 3"
for ci in 1 1.5 2 3; do
python bench.py --sah-ci $ci $NC > gpurun_out/sahci_$ci.json 2> gpurun_out/sahci_$ci.err; python -c "import json;d=json.load(open('gpurun_out/sahci_$ci.json'));print('ci',$ci,d['value'],d['ms_per_step'])"
done
python bench.py --sah-ci 2 --sah-bonus 0.8 $NC > gpurun_out/sahci_2_08.json 2> gpurun_out/sahci_2_08.err; python -c "import json;d=json.load(open('gpurun_out/sahci_2_08.json'));print('ci 2 bonus 0.8',d['value'],d['ms_per_step'])"
echo done
