"""Duck-typed configuration: the Stage-II-relevant defaults of the reference's
support_data/conf/moshpp_conf.yaml (lines 13-31, 34-50, 95-125) without omegaconf (absent here).
``mosh_stageii`` accepts any mapping with attribute/key access (DictConfig, AttrDict, ...)."""
from __future__ import annotations


class AttrDict(dict):
    """Minimal attribute/key mapping standing in for omegaconf's DictConfig (absent here)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(d):
        if isinstance(d, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in d.items()})
        return d


STAGEII_WEIGHTS = dict(stageii_wt_data=400, stageii_wt_velo=2.5, stageii_wt_dmpl=1.0, stageii_wt_expr=1.0,
                       stageii_wt_poseB=1.6, stageii_wt_poseH=1.0, stageii_wt_poseF=1.0, stageii_wt_annealing=2.5)


# Stage I (support_data/conf/moshpp_conf.yaml:99-117, the `smplh` weight block that every surface model type uses by default)
STAGEI_WEIGHTS = dict(stagei_wt_poseH=3.0, stagei_wt_poseF=3., stagei_wt_expr=34., stagei_wt_pose=3., stagei_wt_poseB=3.,
                      stagei_wt_init_finger_left=400.0, stagei_wt_init_finger_right=400.0, stagei_wt_init_finger=400.0,
                      stagei_wt_betas=10., stagei_wt_init=300, stagei_wt_data=75., stagei_wt_surf=10000.,
                      stagei_wt_annealing=[1., .5, .25, .125])


def default_cfg(**over) -> AttrDict:
    """The Stage-II-relevant defaults of support_data/conf/moshpp_conf.yaml (lines 13-31,34-50,95-125)."""
    cfg = AttrDict.wrap({
        'mocap': {'fname': None, 'unit': 'mm', 'rotate': None, 'start_fidx': 0, 'end_fidx': -1, 'ds_rate': 1,
                  'subject_name': 'null', 'multi_subject': False},
        'surface_model': {'type': 'smplx', 'fname': None, 'dmpl_fname': None, 'num_betas': 16,
                          'betas_expr_start_id': 300, 'num_dmpls': 8, 'dof_per_hand': 24, 'num_expressions': 80,
                          'use_hands_mean': True, 'gender': 'neutral'},
        'moshpp': {'pose_body_prior_fname': None, 'pose_hand_prior_fname': None, 'optimize_fingers': False,
                   'optimize_face': False, 'optimize_toes': False, 'optimize_betas': True,
                   'optimize_dynamics': False, 'verbosity': 1},
        'opt_settings': {'weights_type': 'smplh', 'weights': dict(STAGEII_WEIGHTS, **STAGEI_WEIGHTS), 'maxiter': 100,
                         'stagei_lr': 1e-3, 'extra_initial_rigid_adjustment': False},
    })
    for k, v in over.items():
        node = cfg
        parts = k.split('.')
        for q in parts[:-1]:
            node = node[q]
        node[parts[-1]] = v
    return cfg
