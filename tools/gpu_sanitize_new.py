"""Development tool (GPU, for compute-sanitizer): the paths added in round 2 on small cases -- the device-side mocap input
adapter (mosh2_job_upload_markers), the bulk-copy table staging of the Stage-II kernel, the linearise mode + mesh distance of
Stage I (two iterations per minimisation).  Usage: compute-sanitizer --tool memcheck python tools/gpu_sanitize_new.py"""
import copy
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest  # noqa: E402
from moshpp_b200 import chmosh, stagei, synth  # noqa: E402
from moshpp_b200.mocap_interface import MocapSession  # noqa: E402

d = tempfile.mkdtemp()
case = synth.make_case(d, 'C1', **conftest.SMALL['C1'])
out = chmosh.mosh_stageii(case['mocap_fname'], case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'],
                          case['marker_meta'], precision='f32', chunk_len=4, chunk_warmup=4)
print('stage II through the device adapter:', out['stageii_debug_details']['b200']['device_adapter'], out['fullpose'].shape)
case = synth.make_case(d, 'C2', frames=40, n_verts=1500)
cfg = copy.deepcopy(case['cfg'])
cfg.moshpp.optimize_betas = True
cfg.opt_settings.maxiter = 2
frames = MocapSession(case['mocap_fname'], cfg.mocap.unit).markers_asdict()
res = stagei.mosh_stagei([frames[i] for i in (0, 13, 26, 39)], cfg, marker_meta=case['marker_meta'])
print('stage I:', res['stagei_debug_details']['b200'], float(np.abs(res['betas']).max()))
print('OK')
