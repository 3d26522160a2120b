set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02f.txt 2>&1; tail -15 gpurun_out/pytest_gpu_r02f.txt | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29512 bench.py --gpus 2 --config c4 --steps 10 > gpurun_out/r02_bench_c4_n2.json 2> gpurun_out/r02_bench_c4_n2.err; tail -2 gpurun_out/r02_bench_c4_n2.err | cut -c1-400
python bench.py --config c4 --steps 10 > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; tail -2 gpurun_out/r02_bench_c4_n1.err | cut -c1-400
echo done
