set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python profiles/experiments/shard_balance.py > gpurun_out/shard_balance.txt 2>&1; tail -4 gpurun_out/shard_balance.txt
NC="--steps 2 --warmup 3 --no-cpu-baseline --no-parity-check"
for c in c1 c2; do
python bench.py --config $c $NC > gpurun_out/plain_$c.json 2> gpurun_out/plain_$c.err &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_$c python bench.py --config $c $NC > gpurun_out/ncu_r02_$c.log 2>&1
done
echo done
