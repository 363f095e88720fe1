mkdir -p gpurun_out
MOSH2_PROF_LIB=libmosh2_prof.so timeout 120 python tools/gpu_phases.py C2 64 0:0 > gpurun_out/v_phases.txt 2>&1
( timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/v_pytest.log 2>&1
timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
tail -3 gpurun_out/v_pytest.log; head -c 250 gpurun_out/v_bench.json; echo
cut -c1-100 gpurun_out/v_phases.txt | grep -v "warm\.\|minimize"
