set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final2.txt 2>&1; tail -4 gpurun_out/pytest_gpu_final2.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final2.txt 2>&1; tail -1 gpurun_out/smoke_final2.txt
python bench.py > gpurun_out/f2_c3.json 2> gpurun_out/f2_c3.err; python -c "
import json;d=json.load(open('gpurun_out/f2_c3.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity']['words_differ'],d['clocks'])"
echo done
