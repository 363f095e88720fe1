"""Development tool (GPU): mosh2_job_relaunch_chunks must reproduce the rows of a full launch when nothing changes."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moshpp_b200 import chmosh, lib, synth
from moshpp_b200.mocap_interface import MocapSession

for name, frames, hand in (('C4', 400, 'left'), ('C2', 200, None)):
    kw = dict(hand_side=hand) if hand else {}
    case = synth.make_case(tempfile.mkdtemp(), name, frames=frames, **kw)
    pk, opts, _ = chmosh.prepare_stageii(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'])
    mocap = MocapSession(case['mocap_fname'], 'mm')
    obs, vis = mocap.frames_for_labels(case['latent_labels'], range(len(mocap)))
    model = lib.Model(pk, device=0)
    for prec in (lib.MOSH2_F32, lib.MOSH2_F64):
        seq = model.solve(obs, vis, opts, precision=prec)
        job = model.job(frames, opts, chunk_len=14, chunk_warmup=64, warmup_full=-1, precision=prec)
        job.upload(obs, vis); job.launch()
        a = job.download()
        pa, sa = a.pose.copy(), a.status.copy()
        d0 = job.boundary_deltas(a).max(0)
        job.relaunch_chunks([3, 5, 6], 64, -1)
        b = job.download()
        print(name, 'f64' if prec else 'f32', 'chunks', job.num_chunks, 'full launch vs seq', np.abs(pa - seq.pose).max(),
              '| relaunch same W: max |d pose|', np.abs(b.pose - pa).max(), 'status equal', np.array_equal(b.status, sa), '| deltas', d0)
        job.relaunch_chunks([3, 5, 6], 128, -1)
        c = job.download()
        print('    relaunch W=128: vs seq', np.abs(c.pose - seq.pose).max(), 'changed rows', np.nonzero(np.abs(c.pose - pa).max(1) > 0)[0][[0, -1]])
        job.close()
    model.close()
