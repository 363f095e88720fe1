mkdir -p gpurun_out
MOSH2_PROF_LIB=libmosh2_prof.so timeout 120 python tools/gpu_phases.py C2 64 0:0 > gpurun_out/phases_r2.txt 2>&1
( time timeout 1200 python -m pytest tests -x -q -m gpu --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick2.json 2> gpurun_out/bench_quick2.err
tail -25 gpurun_out/pytest_gpu.log; head -c 300 gpurun_out/bench_quick2.json
