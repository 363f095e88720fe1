mkdir -p gpurun_out
MOSH2_PROF_LIB=libmosh2_prof.so timeout 120 python tools/gpu_phases.py C2 64 0:0 > gpurun_out/z_phases.txt 2>&1
cat gpurun_out/z_phases.txt | cut -c1-100
