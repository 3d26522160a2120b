set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L
nproc
( ls -la /usr/lib/x86_64-linux-gnu/ /usr/lib /usr/local/nvidia/lib /usr/lib64 2>/dev/null | grep -i -E "egl|gles|libgl|nvidia|glvnd|gbm" ; ls /usr/share/glvnd/egl_vendor.d /etc/glvnd/egl_vendor.d 2>&1; ldconfig -p | grep -i -E "egl|gles|libGL|opengl" ) > gpurun_out/egl_probe.txt 2>&1
python tests/golden/make_ref_kernel_golden.py gpurun_out/ref_kernel_golden_r02.npz > gpurun_out/golden_r02.txt 2>&1
python tests/golden/ref_kernel_stats.py > gpurun_out/refstats_r02.txt 2>&1
python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_base_c4.json 2> gpurun_out/r02_base_c4.err &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_base_c4 python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02_base_c4.log 2>&1
python bench.py --mode normal --no-jitter --spp 1 --depth 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_base_modea.json 2> gpurun_out/r02_base_modea.err &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_base_modea python bench.py --mode normal --no-jitter --spp 1 --depth 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02_base_modea.log 2>&1
echo done
