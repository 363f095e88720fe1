mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/r02c_pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
timeout 600 python tools/gpu_configs2.py > gpurun_out/r02c_configs.jsonl 2> gpurun_out/r02c_configs.err
MOSH2_PROF_LIB=libmosh2_prof.so timeout 120 python tools/gpu_phases.py C2 64 0:0 > gpurun_out/r02c_phase_clocks_raw.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r02c_launches_bench.log 2>&1
tail -6 gpurun_out/r02c_pytest_gpu.log; head -c 400 gpurun_out/r02c_bench_n1.json; echo; cut -c1-220 gpurun_out/r02c_configs.jsonl
