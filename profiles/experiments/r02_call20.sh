set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export CLPT_LIB=$PWD/clpathtracer_b200/libclpt_busy.so
timeout 600 python profiles/experiments/shard_busy.py > gpurun_out/shard_busy.txt 2> gpurun_out/shard_busy.err; grep -E "^---|claims fill" gpurun_out/shard_busy.err
echo done
