set -x
(timeout 300 python bench.py 2> gpurun_out/s10_bench.err | tail -1) > gpurun_out/s10_bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/s10_launches.csv python bench.py --steps 2 --warmup 1 --frames 2000000 --no-cpu-baseline > gpurun_out/s10_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fe_post|fe_frame4|fe_vad|fe_dc" -s 4 -c 4 -o gpurun_out/s10_fe python tools/fe_sweep.py --once > gpurun_out/s10_fe_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tmat_gemm|tmat_file|tmat_solve|tmat_jacobi" -c 8 -o gpurun_out/s10_tmat python bench.py --no-mfcc --frames 200000 --steps 1 --no-cpu-baseline > gpurun_out/s10_tmat_ncu.log 2>&1
ls -la gpurun_out | tail -12
