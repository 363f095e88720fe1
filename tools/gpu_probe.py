"""Development probe (not part of the product): times the Stage-II kernel on the BASELINE configs for a few
chunk schedules and prints one JSON line per run.  Usage: python tools/gpu_probe.py C2 [frames] [reps]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from moshpp_b200 import chmosh, lib, synth  # noqa: E402
from moshpp_b200.mocap_interface import MocapSession  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else None
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    scheds = [(0, 0), (4, 64), (8, 64), (16, 64), (32, 64), (4, 32)]
    if len(sys.argv) > 4:
        scheds = [tuple(int(x) for x in s.split(':')) for s in sys.argv[4].split(',')]
    d = tempfile.mkdtemp(prefix='mosh_probe_')
    t0 = time.time()
    case = synth.make_case(d, name, frames=frames)
    pk, opts, flags = chmosh.prepare_stageii(case['cfg'], case['markers_latent'], case['latent_labels'],
                                             case['betas'], case['marker_meta'])
    mocap = MocapSession(case['mocap_fname'], 'mm')
    obs, vis = mocap.frames_for_labels(case['latent_labels'], range(len(mocap)))
    F = obs.shape[0]
    print(json.dumps(dict(setup_s=round(time.time() - t0, 2), config=name, frames=F, markers=pk.n_markers,
                          n1=len(pk.free_step1), n2=len(pk.free_step2), kw=pk.kw, na=pk.na)), flush=True)
    model = lib.Model(pk, device=0, library_path=os.environ.get('MOSH2_PROBE_LIB'))   # development: variant builds
    ref = None
    for prec_name, prec in (('f32', lib.MOSH2_F32), ('f64', lib.MOSH2_F64))[:1 if os.environ.get('MOSH2_PROBE_F32_ONLY') else 2]:
        for (L, W) in scheds:
            if prec == lib.MOSH2_F64 and (L, W) not in ((0, 0), (8, 64)):
                continue
            job = model.job(F, opts, chunk_len=L, chunk_warmup=W, precision=prec)
            job.upload(obs, vis)
            ms = []
            for _ in range(reps):
                job.launch()
                job.sync()
                ms.append(job.kernel_ms())
            res = job.download()
            tot = job.totals()
            ok = (res.status & lib.ST_SOLVED) != 0
            if ref is None:
                ref = res.pose.copy()
            d_mk = np.linalg.norm(res.markers_sim - obs, axis=-1)[vis]
            print(json.dumps(dict(prec=prec_name, L=L, W=W, chunks=job.num_chunks, kernel_ms=[round(m, 3) for m in ms],
                                  fps=round(F / (min(ms) * 1e-3), 1), totals=tot,
                                  us_per_build=round(min(ms) * 1e3 / max(1, tot['builds']) * job.num_chunks if L else min(ms) * 1e3 / max(1, tot['builds']), 2),
                                  solved=int(ok.sum()), marker_rms_mm=round(float(np.sqrt((d_mk ** 2).mean()) * 1e3), 4),
                                  max_dpose_vs_first=float(np.abs(res.pose - ref)[ok].max()),
                                  flags_fallback=int(((res.status & lib.ST_GN_FALLBACK) != 0).sum()),
                                  flags_maxiter=int(((res.status & lib.ST_MAXITER) != 0).sum()))), flush=True)
            job.close()
    model.close()


if __name__ == '__main__':
    main()
