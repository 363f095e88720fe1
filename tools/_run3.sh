mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 ) > gpurun_out/pytest_gpu3.log 2>&1
timeout 900 python tools/gpu_stagei.py --oracle > gpurun_out/stagei_full.json 2> gpurun_out/stagei_full.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick3.json 2> gpurun_out/bench_quick3.err
tail -14 gpurun_out/pytest_gpu3.log; cat gpurun_out/stagei_full.json | head -c 1500; tail -3 gpurun_out/stagei_full.err
