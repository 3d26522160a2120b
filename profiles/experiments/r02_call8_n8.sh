set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_c_host.py -m gpu -q -s -k gl_presentation > gpurun_out/pytest_c_host.txt 2>&1; tail -12 gpurun_out/pytest_c_host.txt | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 10 > gpurun_out/r02_bench_c3_n8.json 2> gpurun_out/r02_bench_c3_n8.err; tail -2 gpurun_out/r02_bench_c3_n8.err | cut -c1-400
$TR --master-port 29522 bench.py --gpus 8 --config c4 --steps 20 > gpurun_out/r02_bench_c4_n8.json 2> gpurun_out/r02_bench_c4_n8.err; tail -2 gpurun_out/r02_bench_c4_n8.err | cut -c1-400
$TR --master-port 29523 bench.py --gpus 8 --config c5 --steps 60 > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err; tail -2 gpurun_out/r02_bench_c5_n8.err | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 10 > gpurun_out/r02_bench_c3_n4.json 2> gpurun_out/r02_bench_c3_n4.err
echo done
