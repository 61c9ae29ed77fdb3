set -x
python bench.py --no-cpu-baseline --no-config1 > gpurun_out/pre.json 2> gpurun_out/pre.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --no-cpu-baseline --no-config1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:ozaki_gemm_kernel|narrow_head_kernel|slice_delta_prev|trial_w0_digits|head_fold|narrow_backprop32|gemm_tf32' -c 14 -f -o gpurun_out/prof_r02b python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config1 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/
