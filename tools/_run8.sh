mkdir -p gpurun_out
for v in base pipe t1 all; do
  MOSH2_PROF_LIB=libmosh2_prof_$v.so timeout 120 python tools/gpu_phases.py C2 64 0:0 > gpurun_out/x_phases_$v.txt 2>&1
done
( MOSH2_LIBRARY=$PWD/moshpp_b200/libmosh2_v_all.so timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/x_pytest_all.log 2>&1
for v in base all pipe t1; do
  MOSH2_LIBRARY=$PWD/moshpp_b200/libmosh2_v_$v.so timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/x_bench_$v.json 2> gpurun_out/x_bench_$v.err
done
tail -4 gpurun_out/x_pytest_all.log
for v in base pipe t1 all; do echo $v; grep -E "kernel_ms|bd.T1|gn\.|chunk\(all" gpurun_out/x_phases_$v.txt | cut -c1-90; head -c 200 gpurun_out/x_bench_$v.json; echo; done
