set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_multigpu.py -m gpu -q -s > gpurun_out/pytest_mgpu_r02.txt 2>&1; tail -30 gpurun_out/pytest_mgpu_r02.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench_c3_n2.json 2> gpurun_out/r02_bench_c3_n2.err; tail -2 gpurun_out/r02_bench_c3_n2.err | cut -c1-400
$TR --master-port 29512 bench.py --gpus 2 --config c4 --steps 10 > gpurun_out/r02_bench_c4_n2.json 2> gpurun_out/r02_bench_c4_n2.err; tail -2 gpurun_out/r02_bench_c4_n2.err | cut -c1-400
$TR --master-port 29513 bench.py --gpus 2 --config c5 --steps 60 > gpurun_out/r02_bench_c5_n2.json 2> gpurun_out/r02_bench_c5_n2.err; tail -2 gpurun_out/r02_bench_c5_n2.err | cut -c1-400
$TR --master-port 29514 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n2.json 2> gpurun_out/r02_bench_reference_n2.err
echo done
