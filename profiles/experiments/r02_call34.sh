set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final4.txt 2>&1; tail -1 gpurun_out/smoke_final4.txt
python bench.py > gpurun_out/h_c3.json 2> gpurun_out/h_c3.err; python -c "
import json;d=json.load(open('gpurun_out/h_c3.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity']['words_differ'],d['clocks'],d['gpu_launches'])"
python -m pytest tests/test_gpu_build.py tests/test_c_host.py -m gpu -q > gpurun_out/pytest_final4.txt 2>&1; tail -2 gpurun_out/pytest_final4.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/h_reference.json 2> gpurun_out/h_reference.err; python -c "
import json;d=json.load(open('gpurun_out/h_reference.json'));print(d['value'],d['impl'],d['cpu_baseline']['cores'])"
echo done
