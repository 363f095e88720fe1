mkdir -p gpurun_out
for v in prof_old prof_pf0 prof; do MOSH2_PROF_LIB=libmosh2_$v.so timeout 120 python tools/gpu_phases.py C2 64 0:0 > gpurun_out/phases_$v.txt 2>&1; done
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_parity.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -3 gpurun_out/pytest_parity.log; head -c 600 gpurun_out/bench_quick.json
