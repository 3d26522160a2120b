set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
NC="--no-cpu-baseline --no-parity-check --steps 8 --warmup 3"
CLPT_LIB=$PWD/clpathtracer_b200/libclpt_gb3.so python bench.py $NC > gpurun_out/cd_base.json 2> gpurun_out/cd_base.err; python -c "import json;d=json.load(open('gpurun_out/cd_base.json'));print('base',d['value'],d['ms_per_step'])"
python bench.py $NC > gpurun_out/cd_new.json 2> gpurun_out/cd_new.err; python -c "import json;d=json.load(open('gpurun_out/cd_new.json'));print('countdown',d['value'],d['ms_per_step'])"
CLPT_LIB=$PWD/clpathtracer_b200/libclpt_gb3.so python bench.py --config c4 $NC > gpurun_out/cd_base_c4.json 2> gpurun_out/cd_base_c4.err; python -c "import json;d=json.load(open('gpurun_out/cd_base_c4.json'));print('base c4',d['value'],d['ms_per_step'])"
python bench.py --config c4 $NC > gpurun_out/cd_new_c4.json 2> gpurun_out/cd_new_c4.err; python -c "import json;d=json.load(open('gpurun_out/cd_new_c4.json'));print('countdown c4',d['value'],d['ms_per_step'])"
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_cd.txt 2>&1; tail -5 gpurun_out/pytest_cd.txt
python bench.py --config c5 > gpurun_out/cd_c5.json 2> gpurun_out/cd_c5.err; python -c "import json;d=json.load(open('gpurun_out/cd_c5.json'));print(d['p50_ms'],d['p99_ms'],d['streamed_p50_ms'],d['streamed_p99_ms'],d['parity']['words_differ'])"
echo done
