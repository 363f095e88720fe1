mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/r02_pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 600 python tools/gpu_configs2.py > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err
MOSH2_PROF_LIB=libmosh2_prof.so timeout 120 python tools/gpu_phases.py C2 64 0:0 > gpurun_out/r02_phase_clocks_raw.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mosh2_stageii -c 1 -o gpurun_out/r02_prof python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > gpurun_out/r02_ncu_full.log 2>&1
( timeout 400 compute-sanitizer --tool memcheck python tools/gpu_sanitize_new.py; echo memcheck rc=$? ) > gpurun_out/r02_sanitizer.txt 2>&1
( timeout 400 compute-sanitizer --tool racecheck python tools/gpu_one.py C2 f32; echo racecheck rc=$? ) >> gpurun_out/r02_sanitizer.txt 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -4 gpurun_out/r02_pytest_gpu.log; head -c 300 gpurun_out/r02_bench_n1.json; echo; cut -c1-200 gpurun_out/r02_configs.jsonl; tail -5 gpurun_out/r02_sanitizer.txt; head -c 300 gpurun_out/r02_bench_reference.json
