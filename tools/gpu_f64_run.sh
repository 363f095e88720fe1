mkdir -p gpurun_out
MOSH2_PROF_LIB=libmosh2_prof.so timeout 200 python tools/gpu_phases.py C3 48 0:0 f64 > gpurun_out/p_phases_c3_f64.txt 2>&1
( timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/p_pytest.log 2>&1
timeout 600 python tools/gpu_configs2.py > gpurun_out/p_configs.jsonl 2> gpurun_out/p_configs.err
tail -3 gpurun_out/p_pytest.log; cut -c1-200 gpurun_out/p_configs.jsonl
cut -c1-100 gpurun_out/p_phases_c3_f64.txt | grep -E "kernel_ms|bd\.|chunk\(all|fk\|\|"
