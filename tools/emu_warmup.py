"""Development aid (CPU): deviation of the chunked schedule from the sequential pass, on the host build of the device
source (tests/emu), for warm-up lengths W and numbers of fully solved warm-up frames."""
import sys, os, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ctypes as C
from moshpp_b200 import build, lib, synth
from conftest import dense_obs

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    nv = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
    L = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    handle = C.CDLL(build.build_emu())
    d = tempfile.mkdtemp()
    case = synth.make_case(d, name, frames=frames, n_verts=nv if nv > 0 else None)
    pk, cfg = case['pack'], case['cfg']
    obs, vis = dense_obs(case)
    h = lib.DescHolder(pk)
    opt = lib.make_options(cfg.opt_settings.weights, optimize_fingers=cfg.moshpp.optimize_fingers and pk.finger_hi > pk.finger_lo,
                           optimize_dynamics=cfg.moshpp.optimize_dynamics)
    def solve(chunk_len, W, wf, mode=0, prec=lib.MOSH2_F64):
        F = obs.shape[0]
        res = lib.ResultArrays(F, lib.pack_dims(pk))
        o = np.ascontiguousarray(obs, dtype=np.float64); v8 = np.ascontiguousarray(vis, dtype=np.uint8)
        sched = lib.make_schedule(chunk_len, W, wf, mode)
        t0 = time.time()
        rc = handle.mosh2_emu_solve(C.byref(h.desc), C.byref(opt), F, o.ctypes.data_as(lib._f64p), v8.ctypes.data_as(lib._u8p),
                                    C.byref(sched), prec, C.byref(res.c))
        assert rc == 0
        return res, time.time() - t0
    seq, t = solve(0, 0, -1)
    ok = (seq.status & 1) != 0
    bd = min(pk.body_dof, 66)
    print(f'sequential: {t:.1f}s builds/frame {seq.counters[ok,2].mean():.2f}')
    for W, wf, mode in [(48, -1, 0), (48, 28, 0), (44, 28, 0), (40, 28, 0), (40, 24, 0), (48, 32, 0), (56, 32, 0), (36, 24, 0), (48, 24, 2)]:
        r, t = solve(L, W, wf, mode)
        dp = np.abs(r.pose - seq.pose)[ok]
        dm = np.linalg.norm(r.markers_sim - seq.markers_sim, axis=-1)[ok]
        print(f'L={L} W={W:3d} full={wf:3d} mode={mode:2d} builds={r.counters[:,2].sum()}: body {dp[:, :bd].max():.2e} rad (rms {np.sqrt((dp[:, :bd]**2).mean()):.1e}) '
              f'finger {dp[:, bd:].max() if dp.shape[1] > bd else 0:.2e} trans {np.abs(r.trans - seq.trans)[ok].max()*1e3:.3f} mm '
              f'markers {dm.max()*1e3:.3f} mm  frames>1e-3: {(dp[:, :bd].max(1) > 1e-3).sum()}  {t:.1f}s', flush=True)

main()
