mkdir -p gpurun_out
for v in nopipe pipe; do
  MOSH2_PROF_LIB=libmosh2_prof_$v.so timeout 120 python tools/gpu_phases.py C2 64 0:0 > gpurun_out/y_phases_$v.txt 2>&1
done
for v in nopipe pipe; do
( MOSH2_LIBRARY=$PWD/moshpp_b200/libmosh2_v_$v.so timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py::test_library_is_the_cuda_build ) > gpurun_out/y_pytest_$v.log 2>&1
  MOSH2_LIBRARY=$PWD/moshpp_b200/libmosh2_v_$v.so timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/y_bench_$v.json 2> gpurun_out/y_bench_$v.err
done
for v in nopipe pipe; do echo $v; tail -3 gpurun_out/y_pytest_$v.log; grep -E "kernel_ms|bd.T1|gn\.|chunk\(all|sf\." gpurun_out/y_phases_$v.txt | cut -c1-90; head -c 200 gpurun_out/y_bench_$v.json; echo; done
