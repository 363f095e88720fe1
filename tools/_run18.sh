mkdir -p gpurun_out
MOSH2_PROF_LIB=libmosh2_prof.so timeout 200 python tools/gpu_phases.py C3 48 0:0 f64 > gpurun_out/q_phases_c3_f64.txt 2>&1
MOSH2_PROF_LIB=libmosh2_prof.so timeout 200 python tools/gpu_phases.py C4 64 0:0 f64 > gpurun_out/q_phases_c4_f64.txt 2>&1
MOSH2_PROF_LIB=libmosh2_prof.so timeout 200 python tools/gpu_phases.py C3 48 0:0 f32 > gpurun_out/q_phases_c3_f32.txt 2>&1
for f in c3_f64 c4_f64 c3_f32; do echo $f; cut -c1-100 gpurun_out/q_phases_$f.txt | grep -v "warm\.\|minimize"; done
