set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 > gpurun_out/f_c3_n8.json 2> gpurun_out/f_c3_n8.err; tail -2 gpurun_out/f_c3_n8.err | cut -c1-300
$TR --master-port 29522 bench.py --gpus 8 --config c4 > gpurun_out/f_c4_n8.json 2> gpurun_out/f_c4_n8.err
$TR --master-port 29523 bench.py --gpus 8 --config c5 > gpurun_out/f_c5_n8.json 2> gpurun_out/f_c5_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 > gpurun_out/f_c3_n4.json 2> gpurun_out/f_c3_n4.err
python bench.py > gpurun_out/f_c3_n1_samebox.json 2> gpurun_out/f_c3_n1_samebox.err
for f in c3_n8 c4_n8 c5_n8 c3_n4 c3_n1_samebox; do python -c "
import json;d=json.load(open('gpurun_out/f_$f.json'));p=d.get('parity') or {};print('$f',d.get('value'),d.get('ms_per_step'),(d.get('e2e') or {}).get('value'),p.get('words_differ'),d.get('p99_ms'),d.get('per_rank_kernel_ms'))" | cut -c1-600; done
echo done
