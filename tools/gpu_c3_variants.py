import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_long_golden import LONG
from moshpp_b200 import chmosh, synth
key = sys.argv[1] if len(sys.argv) > 1 else 'C3F'
name, kw = LONG[key]
case = synth.make_case(tempfile.mkdtemp(), name, **kw)
g = np.load(os.path.join(ROOT, 'tests', 'golden', f'long_{key}.npz'))
variants = [dict(), dict(boundary_tol=(3e-5, 3e-4, 3e-6, 3e-4)), dict(boundary_tol=(1e-5, 1e-4, 1e-6, 1e-4)),
            dict(chunk_warmup=96, warmup_full=80, boundary_tol=(3e-5, 3e-4, 3e-6, 3e-4)), dict(precision='f32', boundary_tol=(3e-5, 3e-4, 3e-6, 3e-4))]
for v in variants:
    out = chmosh.mosh_stageii(case['mocap_fname'], case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'], **v)
    b = out['stageii_debug_details']['b200']
    dp = np.abs(b['pose_reduced'] - g['pose'])
    body = dp[:, :66].max(1)
    bc = b['boundary_check']
    print(v, '|', b['precision'], 'kernel', round(b['kernel_ms'], 1), 'ms rounds', bc['rounds'], bc['repaired_chunks'], 'unverified', bc['unverified_chunks'],
          '| body max %.2e rms %.1e over1e-3: %d (%.2f%%) finger over: %d trans over: %d' % (body.max(), np.sqrt((body**2).mean()), (body > 1e-3).sum(), 100 * (body > 1e-3).mean(),
          (dp[:, 66:].max(1) > 1e-2).sum(), (np.abs(out['trans'] - g['trans']).max(1) > 1e-4).sum()), flush=True)
