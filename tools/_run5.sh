mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/r02_pytest_gpu.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 900 python tools/gpu_configs2.py > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r02_launches_bench.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log; head -c 300 gpurun_out/r02_bench_n1.json; echo; cat gpurun_out/r02_configs.jsonl | cut -c1-200
