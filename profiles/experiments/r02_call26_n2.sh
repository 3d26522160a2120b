set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/pytest_multigpu_final.txt 2>&1; tail -4 gpurun_out/pytest_multigpu_final.txt | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 > gpurun_out/f_c3_n2.json 2> gpurun_out/f_c3_n2.err; tail -2 gpurun_out/f_c3_n2.err | cut -c1-300
$TR --master-port 29512 bench.py --gpus 2 --config c4 > gpurun_out/f_c4_n2.json 2> gpurun_out/f_c4_n2.err
$TR --master-port 29513 bench.py --gpus 2 --config c5 > gpurun_out/f_c5_n2.json 2> gpurun_out/f_c5_n2.err
$TR --master-port 29514 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/f_reference_n2.json 2> gpurun_out/f_reference_n2.err
for f in c3_n2 c4_n2 c5_n2 reference_n2; do python -c "
import json;d=json.load(open('gpurun_out/f_$f.json'));print('$f',d.get('value'),d.get('ms_per_step'),(d.get('e2e') or {}).get('value'),(d.get('parity') or {}),d.get('p99_ms'))" | cut -c1-600; done
echo done
