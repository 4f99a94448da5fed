set -x
python bench.py > gpurun_out/f_bench_fp32.json 2> gpurun_out/f_bench_fp32.err
SRB_BENCH_PRECISION=64 python bench.py --no-cpu-baseline > gpurun_out/f_bench_fp64.json 2> gpurun_out/f_bench_fp64.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:srb_step_kernel -s 10 -c 2 -o gpurun_out/f_step_f32 python tools/prof_step.py 32 8 4096 14 > gpurun_out/f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:srb_step_kernel -s 10 -c 2 -o gpurun_out/f_step_f64 python tools/prof_step.py 64 8 4096 14 > gpurun_out/f_ncu2.log 2>&1
tail -c 600 gpurun_out/f_bench_fp32.json
