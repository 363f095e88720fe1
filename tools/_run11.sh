mkdir -p gpurun_out
for v in prof prof_quiet; do
MOSH2_PROF_LIB=libmosh2_$v.so timeout 120 python tools/gpu_phases.py C2 64 0:0 > gpurun_out/w_phases_$v.txt 2>&1
done
( timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/w_pytest.log 2>&1
timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err
tail -3 gpurun_out/w_pytest.log; head -c 250 gpurun_out/w_bench.json; echo
for v in prof prof_quiet; do cut -c1-100 gpurun_out/w_phases_$v.txt | grep -v "warm\.\|minimize"; done
