#!/bin/bash
# Run on a B200: collects everything scripts/refresh_profiles.sh turns into profiles/.  Under gpurun use
# scripts/gpu_profile_part.sh A and B instead: the four .ncu-rep files together exceed the 64 MiB gpurun brings back.
mkdir -p gpurun_out
R=${ROUND:-r02}
python bench.py > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${R}_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --strong-res 0 > gpurun_out/launches_bench.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:integrate_kernel -c 3 --csv --log-file gpurun_out/traffic_paged.csv \
    python bench.py --steps 1 --warmup 1 --no-render --no-cpu-baseline > gpurun_out/traffic_paged.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 1 -c 1 -f -o gpurun_out/prof_integrate_final \
    python scripts/profile_target.py integrate > gpurun_out/prof_integrate_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 1 -c 1 -f -o gpurun_out/prof_integrate_paged \
    python scripts/profile_target.py paged > gpurun_out/prof_integrate_paged.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/prof_render \
    python scripts/profile_target.py render 256 > gpurun_out/prof_render.log 2>&1
MK_LONG_EXCLUSIVE=0 ncu --set full --clock-control none --import-source on -k regex:render_pipeline_kernel -s 1 -c 1 -f -o gpurun_out/prof_render_long \
    python scripts/profile_target.py long > gpurun_out/prof_render_long.log 2>&1
python scripts/parity_report.py > gpurun_out/parity_report.log 2>&1
ls -la gpurun_out | tail -20
