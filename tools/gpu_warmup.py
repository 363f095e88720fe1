"""Development tool (GPU): deviation of the chunked (parallel-in-time) schedule from the sequential pass as a function
of the warm-up length, per precision, split into body pose / finger coefficients / translation / simulated markers.

    python tools/gpu_warmup.py C2 500 16,24,32,48,64
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moshpp_b200 import chmosh, lib, synth  # noqa: E402
from moshpp_b200.mocap_interface import MocapSession  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    warm = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else '16,24,32,48,64').split(',')]
    d = tempfile.mkdtemp(prefix='mosh_warm_')
    case = synth.make_case(d, name, frames=frames)
    pk, opts, _ = chmosh.prepare_stageii(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'])
    mocap = MocapSession(case['mocap_fname'], 'mm')
    obs, vis = mocap.frames_for_labels(case['latent_labels'], range(len(mocap)))
    F = obs.shape[0]
    bd = min(pk.body_dof, 66)
    model = lib.Model(pk, device=0)
    L = chmosh.auto_chunk_len(F)
    seq = {}
    for prec_name, prec in (('f64', lib.MOSH2_F64), ('f32', lib.MOSH2_F32)):
        for W in [None] + warm:
            job = model.job(F, opts, chunk_len=0 if W is None else L, chunk_warmup=W or 0, precision=prec)
            job.upload(obs, vis)
            job.launch()
            res = job.download()
            ms = job.kernel_ms()
            job.close()
            if W is None:
                seq[prec_name] = res
                if prec_name == 'f32':
                    r64 = seq['f64']
                    print(json.dumps(dict(what='f32 sequential vs f64 sequential', body=float(np.abs(res.pose - r64.pose)[:, :bd].max()),
                                          fingers=float(np.abs(res.pose - r64.pose)[:, bd:].max()) if res.pose.shape[1] > bd else 0.0,
                                          trans_mm=float(np.abs(res.trans - r64.trans).max() * 1e3),
                                          markers_mm=float(np.abs(res.markers_sim - r64.markers_sim).max() * 1e3))), flush=True)
                continue
            ref = seq[prec_name]
            dp = np.abs(res.pose - ref.pose)
            print(json.dumps(dict(prec=prec_name, L=L, W=W, kernel_ms=round(ms, 2), fps=round(F / (ms * 1e-3)),
                                  body=float(dp[:, :bd].max()), body_rms=float(np.sqrt((dp[:, :bd] ** 2).mean())),
                                  fingers=float(dp[:, bd:].max()) if dp.shape[1] > bd else 0.0,
                                  trans_mm=float(np.abs(res.trans - ref.trans).max() * 1e3),
                                  markers_mm=float(np.abs(res.markers_sim - ref.markers_sim).max() * 1e3))), flush=True)
    model.close()


if __name__ == '__main__':
    main()
