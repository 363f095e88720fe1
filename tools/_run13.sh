mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/u_pytest.log 2>&1
( MOSH2_LIBRARY=$PWD/moshpp_b200/libmosh2_v_gold.so timeout 900 python -m pytest tests/test_golden.py tests/test_gpu_parity.py -q -m gpu --deselect tests/test_gpu_parity.py::test_library_is_the_cuda_build ) > gpurun_out/u_pytest_gold.log 2>&1
timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
tail -4 gpurun_out/u_pytest.log; tail -4 gpurun_out/u_pytest_gold.log; head -c 250 gpurun_out/u_bench.json; echo
