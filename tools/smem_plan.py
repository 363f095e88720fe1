"""Development tool (CPU): shared-memory footprint of the Stage-II workspace for the BASELINE configurations, from the
same carve() the kernel uses (through the test-only host build).  Usage: python tools/smem_plan.py [C1 C2 ...]"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moshpp_b200 import build, chmosh, lib, synth  # noqa: E402


def main():
    os.environ['MOSH2_EMU_PLAN'] = '1'
    handle = C.CDLL(build.build_emu())
    d = tempfile.mkdtemp(prefix='mosh_plan_')
    for name in (sys.argv[1:] or ['C1', 'C2', 'C3', 'C4']):
        case = synth.make_case(d, name, frames=1)
        pk, opts, _ = chmosh.prepare_stageii(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'])
        h = lib.DescHolder(pk)
        M = pk.n_markers
        obs = np.zeros((1, M, 3))
        vis = np.zeros((1, M), dtype=np.uint8)           # nothing visible: the frame is skipped, only the plan prints
        res = lib.ResultArrays(1, lib.pack_dims(pk))
        print(name, 'markers', M, 'n1', len(pk.free_step1), 'n2', len(pk.free_step2), flush=True)
        for prec in (lib.MOSH2_F32, lib.MOSH2_F64):
            handle.mosh2_emu_solve(C.byref(h.desc), C.byref(opts), 1, obs.ctypes.data_as(lib._f64p),
                                   vis.ctypes.data_as(lib._u8p), 0, 0, prec, C.byref(res.c))


if __name__ == '__main__':
    main()
