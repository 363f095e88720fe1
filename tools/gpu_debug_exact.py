import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_long_golden import LONG
from moshpp_b200 import chmosh, lib, synth
from moshpp_b200.mocap_interface import MocapSession
name, kw = LONG['C4L']
case = synth.make_case(tempfile.mkdtemp(), name, **kw)
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'long_C4L.npz'))
pk, opts, _ = chmosh.prepare_stageii(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'])
mocap = MocapSession(case['mocap_fname'], 'mm')
obs, vis = mocap.frames_for_labels(case['latent_labels'], range(len(mocap)))
model = lib.Model(pk, device=0)
fid = g['frame_ids']
def err(res): return np.abs(res.pose[fid] - g['pose']).max(1)
job = model.job(2000, opts, chunk_len=14, chunk_warmup=256, warmup_full=-1, precision=lib.MOSH2_F64)
job.upload(obs, vis); job.launch(); a = job.download()
e = err(a); print('launch W=256: max', e.max(), 'frames>1e-3', (e > 1e-3).sum(), 'ms', job.kernel_ms())
d = job.boundary_deltas(a); bad = np.nonzero((d > np.array(chmosh.BOUNDARY_TOL['exact'])[None]).any(1))[0]
print('bad chunks', len(bad), bad[:10], 'delta max', d.max(0))
for W in (512, 1024):
    job.relaunch_chunks(bad, W, W); b = job.download()
    e = err(b); print(f'relaunch {len(bad)} chunks W={W}: max', e.max(), 'frames>1e-3', (e > 1e-3).sum(), 'first bad frames', np.nonzero(e > 1e-3)[0][:10], 'ms', job.kernel_ms(),
                      'status or', np.bitwise_or.reduce(b.status))
    d = job.boundary_deltas(b); bad = np.nonzero((d > np.array(chmosh.BOUNDARY_TOL['exact'])[None]).any(1))[0]
    print('   bad chunks now', len(bad), 'delta max', d.max(0))
job.close()
seq = model.solve(obs, vis, opts, precision=lib.MOSH2_F64)
print('sequential f64: max', err(seq).max())
