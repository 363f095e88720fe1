import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import conftest
from moshpp_b200 import synth, lib
import tempfile
name = sys.argv[1] if len(sys.argv) > 1 else 'C4'
case = synth.make_case(tempfile.mkdtemp(), name, **conftest.SMALL[name])
ref = conftest.gpu_solve(case, precision='f64')
for thr in (None, '256', '128', '512'):
    if thr: os.environ['MOSH2_DEV_THREADS'] = thr
    for rep in range(2):
        r = conftest.gpu_solve(case, precision='f32')
        print(name, 'threads', thr, 'rep', rep, 'dtrans', float(np.abs(r.trans - ref.trans).max()), 'dpose', float(np.abs(r.pose - ref.pose).max()),
              'builds', r.counters[:, 2].tolist(), 'flags', (r.status & 48).tolist())
print('f64 builds', ref.counters[:, 2].tolist())
