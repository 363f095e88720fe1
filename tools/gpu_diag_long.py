"""Development tool (GPU): where does a long run leave the sequential float64 oracle?  For one of the long fixtures
(tests/golden/long_*.npz) runs the kernel in several modes and lists the frames over tolerance.
Usage: python tools/gpu_diag_long.py NS [modes...]   modes: f32seq f64chunk f32chunk f32chunk_full"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import tempfile
from make_long_golden import LONG
from moshpp_b200 import chmosh, lib, synth
from moshpp_b200.mocap_interface import MocapSession


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else 'NS'
    modes = sys.argv[2:] or ['f32seq', 'f64chunk', 'f32chunk', 'f32chunk_full']
    name, kw = LONG[key]
    case = synth.make_case(tempfile.mkdtemp(), name, **kw)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', f'long_{key}.npz'))
    pk, opts, _ = chmosh.prepare_stageii(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'])
    mocap = MocapSession(case['mocap_fname'], 'mm')
    obs, vis = mocap.frames_for_labels(case['latent_labels'], range(len(mocap)))
    F = obs.shape[0]
    L = chmosh.plan_chunk_len([F])
    model = lib.Model(pk, device=0)
    fid = g['frame_ids']
    bd = min(pk.body_dof, 66)
    results = {}
    for mode in modes:
        prec = lib.MOSH2_F64 if mode.startswith('f64') else lib.MOSH2_F32
        if ':' in mode:                                  # f32:L:W:WF  (L = 0: planned chunk length)
            _, a, b, c = mode.split(':')
            cl, W, WF = (int(a) or L), int(b), int(c)
        else:
            cl, W, WF = (0, 0, -1) if mode.endswith('seq') else (L, 48, -1 if mode.endswith('full') else 32)
        res = model.solve(obs, vis, opts, chunk_len=cl, chunk_warmup=W, warmup_full=WF, precision=prec)
        results[mode] = res
        dp = np.abs(res.pose[fid] - g['pose'])
        body = dp[:, :bd].max(1)
        tolb = 5e-3 if pk.model_type == 'mano' else 1e-3
        bad = np.nonzero(body > tolb)[0]
        print(f'== {key} {mode}: chunk_len {cl} W {W} full {WF}: body max {body.max():.2e} at frame {fid[body.argmax()]}, '
              f'finger max {dp[:, bd:].max():.2e}, trans max {np.abs(res.trans[fid] - g["trans"]).max():.2e}, '
              f'frames > tol: {len(bad)}, data SSE rel max {np.abs(res.errs[fid, 0] / g["err_data"] - 1).max():.2e} (frames > 1 %: {(np.abs(res.errs[fid, 0] / g["err_data"] - 1) > 1e-2).sum()}), builds (emitted) {res.counters[fid, 2].sum()} vs oracle {int(g["j_evals"])}')
        if len(bad):
            runs, start = [], bad[0]
            for a, b in zip(bad[:-1], bad[1:]):
                if b != a + 1:
                    runs.append((start, a)); start = b
            runs.append((start, bad[-1]))
            for a, b in runs[:12]:
                print(f'     frames {fid[a]}..{fid[b]} (chunk {fid[a] // max(cl, 1) if cl else 0}, offset in chunk {fid[a] % cl if cl else fid[a]}): '
                      f'max {body[a:b + 1].max():.2e}; nvis {vis[fid[a]].sum()}; builds here {res.counters[fid[a]:fid[b] + 1, 2].tolist()[:12]}')
    if 'f32seq' in results and 'f32chunk' in results:
        d = np.abs(results['f32seq'].pose[fid] - results['f32chunk'].pose[fid])[:, :bd].max(1)
        print(f'f32chunk vs f32seq: max {d.max():.2e} at {fid[d.argmax()]}, frames > 1e-3: {(d > 1e-3).sum()}')
    model.close()


main()
