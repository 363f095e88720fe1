mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 ) > gpurun_out/r02b_pytest_gpu.log 2>&1
timeout 900 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
tail -8 gpurun_out/r02b_pytest_gpu.log; head -c 1500 gpurun_out/r02b_bench_n1.json; echo; tail -3 gpurun_out/r02b_bench_n1.err
