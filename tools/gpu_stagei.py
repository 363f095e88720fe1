"""Development tool (GPU): Stage I at BASELINE size -- twelve frames of the 4000-frame SMPL-H sequence, 53 markers, 16 shape
coefficients -- through moshpp_b200.stagei.mosh_stagei; with --oracle also the float64 oracle on the host cores (parity + time).
Usage: python tools/gpu_stagei.py [--oracle] [--frames 12]"""
import argparse
import copy
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moshpp_b200 import stagei, synth  # noqa: E402
from moshpp_b200.mocap_interface import MocapSession  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--oracle', action='store_true')
    ap.add_argument('--frames', type=int, default=12)
    ap.add_argument('--config', default='C5')
    a = ap.parse_args()
    d = tempfile.mkdtemp(prefix='mosh_stagei_')
    case = synth.make_case(d, a.config, frames=480)
    cfg = copy.deepcopy(case['cfg'])
    cfg.moshpp.optimize_betas = True
    mocap = MocapSession(case['mocap_fname'], cfg.mocap.unit)
    frames = mocap.markers_asdict()
    pick = np.linspace(0, len(frames) - 1, a.frames).astype(int)
    frames = [frames[i] for i in pick]
    t0 = time.perf_counter()
    out = stagei.mosh_stagei(frames, cfg, marker_meta=case['marker_meta'])
    dt = time.perf_counter() - t0
    st = out['stagei_debug_details']['b200']
    nb = cfg.surface_model.num_betas
    line = {'workload': f'Stage I: {a.frames} frames, {len(out["latent_labels"])} markers, {nb} betas, model {cfg.surface_model.type}',
            'seconds': dt, 'stats': st, 'errs': out['stagei_debug_details']['stagei_errs'],
            'betas_err_vs_truth': float(np.abs(out['betas'][:nb] - case['betas'][:nb]).max()),
            'latent_err_vs_truth_mm': float(1e3 * np.abs(out['markers_latent'] - case['markers_latent']).max())}
    if a.oracle:
        from oracle import stagei as ostagei
        t0 = time.perf_counter()
        ref = ostagei.mosh_stagei(frames, cfg, marker_meta=case['marker_meta'])
        line['oracle_seconds'] = time.perf_counter() - t0
        line['oracle_stats'] = ref['stagei_debug_details']['oracle_stats']
        line['d_betas'] = float(np.abs(out['betas'] - ref['betas']).max())
        line['d_latent'] = float(np.abs(out['markers_latent'] - ref['markers_latent']).max())
        line['d_pose'] = float(max(np.abs(p - q).max() for p, q in zip(out['stagei_debug_details']['opt_models_pose'],
                                                                        ref['stagei_debug_details']['opt_models_pose'])))
    print(json.dumps(line))


if __name__ == '__main__':
    main()
