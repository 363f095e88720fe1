"""Stage-II result writers: the ``*_stageii.pkl`` merge and the AMASS ``.npz`` layout.

Restates, for the keys Stage II produces, ``MoSh.mosh_stageii``'s post-processing
(src/moshpp/mosh_head.py:289-295), ``MoSh.load_as_amass_npz`` (mosh_head.py:444-541) and
``turn_fullpose_into_parts`` (tools/run_tools.py:70-85).  omegaconf is absent here, so ``cfg`` is any nested
mapping with item access.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, Optional, Union

import numpy as np


def turn_fullpose_into_parts(fullpose: np.ndarray, surface_model_type: str) -> Dict[str, np.ndarray]:
    """tools/run_tools.py:70-85."""
    res = {'root_orient': fullpose[:, :3]}
    if 'smpl' in surface_model_type:
        res['pose_body'] = fullpose[:, 3:66]
    elif any(t in surface_model_type for t in ('animal', 'object')):
        res['pose_body'] = fullpose[:, 3:]
    if 'smplh' in surface_model_type:
        res['pose_hand'] = fullpose[:, 66:]
    elif 'smplx' in surface_model_type:
        res['pose_hand'] = fullpose[:, 75:]
        res['pose_jaw'] = fullpose[:, 66:69]
        res['pose_eye'] = fullpose[:, 69:75]
    elif 'mano' in surface_model_type:
        res['pose_hand'] = fullpose[:, 3:]
    return res


def _to_plain(cfg):
    if isinstance(cfg, dict):
        return {k: _to_plain(v) for k, v in cfg.items()}
    return cfg


def merge_stageii(stageii_data: dict, stagei_data: dict, cfg, elapsed_time: float,
                  stageii_fname: Optional[str] = None) -> dict:
    """What ``MoSh.mosh_stageii`` does with the solver's return value (mosh_head.py:289-295)."""
    stageii_data.update(stagei_data)
    stageii_data['stageii_debug_details']['stageii_elapsed_time'] = elapsed_time
    stageii_data['stageii_debug_details']['cfg'] = _to_plain(cfg)
    if stageii_fname:
        os.makedirs(os.path.dirname(os.path.abspath(stageii_fname)), exist_ok=True)
        with open(stageii_fname, 'wb') as f:
            pickle.dump(stageii_data, f)
    return stageii_data


STAGEI_NPZ_KEYS = ('gender', 'surface_model_type', 'markers_latent', 'latent_labels', 'markers_latent_vids', 'betas', 'v_template')


def load_as_amass_npz(stageii_pkl_data_or_fname: Union[dict, str], stageii_npz_fname: Optional[str] = None,
                      stagei_npz_fname: Optional[str] = None, include_markers: bool = False,
                      include_extra_details: bool = False) -> dict:
    """AMASS npz dictionary of a merged Stage-II result, same arguments as the reference (mosh_head.py:444-541): with
    ``stageii_npz_fname`` the sequence file is written and next to it (or at ``stagei_npz_fname``) the per-subject
    ``<gender>_stagei.npz`` with the Stage-I keys, each only if it does not exist yet.  The reference slices ``dmpls``
    along frames by mistake (``[:num_dmpls]``, mosh_head.py:499); here the coefficient
    axis is sliced, which is what AMASS files contain."""
    if isinstance(stageii_pkl_data_or_fname, dict):
        d = stageii_pkl_data_or_fname
    else:
        with open(stageii_pkl_data_or_fname, 'rb') as f:
            d = pickle.load(f)
    dbg = d['stageii_debug_details']
    cfg = dbg['cfg']
    sm, mp = cfg['surface_model'], cfg['moshpp']
    out = {
        'gender': sm.get('gender', 'neutral'),
        'surface_model_type': sm['type'],
        'mocap_frame_rate': dbg['mocap_frame_rate'],
        'mocap_time_length': dbg['mocap_time_length'],
        'markers_latent': d['markers_latent'],
        'latent_labels': d['latent_labels'],
        'markers_latent_vids': d.get('markers_latent_vids'),
        'trans': d['trans'],
        'poses': d['fullpose'],
    }
    if include_extra_details:
        out['surface_model_fname'] = sm.get('fname')
    if 'v_template' in d.get('stagei_debug_details', {}):
        out['v_template'] = d['stagei_debug_details']['v_template']
    if mp.get('optimize_betas', True):
        out['betas'] = np.asarray(d['betas'])[:sm['num_betas']]
        out['num_betas'] = sm['num_betas']
    if mp.get('optimize_dynamics', False):
        out['dmpls'] = np.asarray(d['dmpls'])[:, :sm['num_dmpls']]
        out['num_dmpls'] = sm['num_dmpls']
    if mp.get('optimize_face', False):
        out['expression'] = np.asarray(d['expression'])[:, :sm['num_expressions']]
        out['num_expressions'] = sm['num_expressions']
    out.update(turn_fullpose_into_parts(np.asarray(d['fullpose']), sm['type']))
    if include_markers:
        out['markers'] = dbg['markers_orig']
        out['labels'] = dbg['labels_orig']
        out['markers_obs'] = np.array(dbg['markers_obs'], dtype=object)
        out['labels_obs'] = np.array(dbg['labels_obs'], dtype=object)
        out['markers_sim'] = np.array(dbg['markers_sim'], dtype=object)
        out['marker_meta'] = d.get('marker_meta')
        out['num_markers'] = np.asarray(dbg['markers_orig']).shape[1]
    if stageii_npz_fname:
        if not os.path.exists(stageii_npz_fname):
            os.makedirs(os.path.dirname(os.path.abspath(stageii_npz_fname)), exist_ok=True)
            np.savez(stageii_npz_fname, **{k: v for k, v in out.items() if v is not None})
        if stagei_npz_fname is None:
            stagei_npz_fname = os.path.join(os.path.dirname(os.path.abspath(stageii_npz_fname)), f"{out['gender']}_stagei.npz")
        if not os.path.exists(stagei_npz_fname):
            os.makedirs(os.path.dirname(os.path.abspath(stagei_npz_fname)), exist_ok=True)
            np.savez(stagei_npz_fname, **{k: v for k, v in out.items() if k in STAGEI_NPZ_KEYS and v is not None})
    return out
