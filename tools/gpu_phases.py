"""Development tool: phase clock breakdown of the Stage-II kernel (needs moshpp_b200/libmosh2_prof.so built with
-DMOSH2_PROFILE).  Usage: python tools/gpu_phases.py C2 [frames] [L:W] [f32|f64]"""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moshpp_b200 import chmosh, lib, synth  # noqa: E402
from moshpp_b200.mocap_interface import MocapSession  # noqa: E402

NAMES = ['ev.fullpose', 'ev.rodrigues', 'ev.fk||blend', 'ev.skin+prior', 'ev.markers', 'ev.reduce', 'bd.pre', 'bd.T1',
         'bd.T2', 'bd.T3', 'bd.closed', 'gn.init', 'gn.w0 panel+update', 'gn.w0 diag block', 'gn.w0 barrier wait', 'gn.solves', 'minimize(all)',
         'chunk(all)', 'ev.fk alone', 'ev.prior alone', 'sf.stage_setup', 'sf.accept logic', 'sf.symv+reduce (pre GN)', 'sf.step+symv (post GN)', 'sf.output', 'sf.other']


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    L, W = (int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else '0:0').split(':'))
    prec = lib.MOSH2_F64 if (len(sys.argv) > 4 and sys.argv[4] == 'f64') else lib.MOSH2_F32
    d = tempfile.mkdtemp(prefix='mosh_phase_')
    case = synth.make_case(d, name, frames=frames)
    pk, opts, _ = chmosh.prepare_stageii(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'])
    mocap = MocapSession(case['mocap_fname'], 'mm')
    obs, vis = mocap.frames_for_labels(case['latent_labels'], range(len(mocap)))
    path = os.path.join(ROOT, 'moshpp_b200', os.environ.get('MOSH2_PROF_LIB', 'libmosh2_prof.so'))
    model = lib.Model(pk, device=0, library_path=path)
    job = model.job(obs.shape[0], opts, chunk_len=L, chunk_warmup=W, precision=prec)
    job.upload(obs, vis)
    job.launch(); job.sync()
    job.launch(); job.sync()
    ms = job.kernel_ms()
    tot = job.totals()
    clk = np.zeros(32, dtype=np.int64)
    model.lib.mosh2_dev_phase_clocks.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    model.lib.mosh2_dev_phase_clocks(job.handle, clk.ctypes.data_as(C.POINTER(C.c_longlong)))
    nb = max(1, tot['builds'])
    print(json.dumps(dict(kernel_ms=ms, totals=tot, chunks=job.num_chunks)))
    chunk = clk[17]
    for i, n in enumerate(NAMES):
        print(f'{n:16s} {clk[i]/nb:10.0f} cycles/build  {100*clk[i]/max(1,chunk):5.1f}% of chunk time')


if __name__ == '__main__':
    main()
