"""Mocap input adapter with the reference's label clean-up, units and visibility rule.

Restates tools/mocap_interface.py:87-162 (``read_mocap``) and :165-295 (``MocapSession``) for the
formats that can be read here: .npz, .pkl, .mat (scipy) and .c3d (through ``c3d_io`` because ezc3d
is not installable).  The dense ``frames_for_labels`` view replaces the per-frame dictionaries of
``markers_asdict`` (:254-273) that the Stage-II loop consumes at chmosh.py:582-594.
"""
from __future__ import annotations

import pickle
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np


def rotate_points_xyz(points: np.ndarray, rxyz_deg: Sequence[float]) -> np.ndarray:
    """F x N x 3 points rotated by Rz Ry Rx (degrees), as human_body_prior's helper of the same name."""
    ax, ay, az = np.radians(np.asarray(rxyz_deg, dtype=np.float64).ravel()[:3])
    rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    R = rz @ ry @ rx
    return np.einsum('cd,fnd->fnc', R, points)


def read_mocap(mocap_fname: str) -> Dict:
    """tools/mocap_interface.py:87-162."""
    labels = None
    frame_rate = None
    mocap_fname = str(mocap_fname)
    if mocap_fname.endswith('.mat'):
        import scipy.io
        md = scipy.io.loadmat(mocap_fname)
        markers = None
        for key in ('MoCaps', 'Markers'):
            if key in md:
                markers = md[key]
        if markers is None:
            raise ValueError("The .mat file do not have the expected field for marker data! "
                             "Expected fields are ['MoCaps', 'Markers']")
        if 'Labels' in md:
            labels = np.vstack(md['Labels'][0]).ravel().tolist()
    elif mocap_fname.endswith('.pkl'):
        with open(mocap_fname, 'rb') as f:
            md = pickle.load(f, encoding='latin-1')
        markers = md['markers']
        if 'required_parameters' in md:
            frame_rate = md['required_parameters']['frame_rate']
        elif 'frame_rate' in md:
            frame_rate = md['frame_rate']
        labels = md.get('labels', False)
        if isinstance(labels, np.ndarray):
            labels = labels.tolist()
        labels = [f'*{i}' if isinstance(l, np.ndarray) else l for i, l in enumerate(labels)] if labels else None
    elif mocap_fname.endswith('.c3d'):
        from .c3d_io import read_c3d
        markers, labels, frame_rate = read_c3d(mocap_fname)
        if len(labels) < markers.shape[1]:
            labels = labels + [f'*{len(labels) + i:d}' for i in range(markers.shape[1] - len(labels))]
    elif mocap_fname.endswith('.npz'):
        md = np.load(mocap_fname, allow_pickle=True)
        markers = md['markers']
        if 'frame_rate' in md.files:
            frame_rate = float(md['frame_rate'])
        elif 'required_parameters' in md.files:
            rp = md['required_parameters'].item()
            if 'frame_rate' in rp:
                frame_rate = rp['frame_rate']
        labels = md['labels'].tolist() if 'labels' in md.files else None
    else:
        raise ValueError(f"Error! Could not recognize file format for {mocap_fname}")

    markers = np.array(markers, dtype=np.float64)
    if labels is None:
        labels = [f'*{i}' for i in range(markers.shape[1])]
    elif len(labels) < markers.shape[1]:
        labels = list(labels) + [f'*{i}' for i in range(markers.shape[1] - len(labels))]
    labels = [l.decode() if isinstance(l, bytes) else str(l) for l in labels]

    subject_mask, subject_id_map = [], {}
    for l in labels:
        name = l.split(':')[0] if ':' in l else 'null'
        if name not in subject_id_map:
            subject_id_map[name] = len(subject_id_map)
        subject_mask.append(subject_id_map[name])
    subject_mask = {n: np.array([i == sid for i in subject_mask], dtype=bool) for n, sid in subject_id_map.items()}
    return {'markers': markers, 'labels': labels, 'frame_rate': frame_rate, 'subject_mask': subject_mask}


class MocapSession:
    """tools/mocap_interface.py:165-295 (reader side)."""

    def __init__(self, mocap_fname, mocap_unit: str, mocap_rotate=None, exclude_markers: List[str] = None,
                 only_subjects: List[str] = None, only_markers: List[str] = None, labels_map: dict = None,
                 ignore_stared_labels: bool = True, remove_label_before_colon: bool = True):
        scale = {'mm': 1000., 'cm': 100., 'm': 1.}[mocap_unit]
        self.mocap_fname = mocap_fname
        self.read_status = False
        if only_subjects:
            assert isinstance(only_subjects, list), ValueError(
                f'attribute only_subjects should be a list of strings as subject names: {only_subjects}')
        rd = read_mocap(mocap_fname)
        labels = [l.replace(' ', '') for l in rd['labels']]
        if remove_label_before_colon:
            labels = [l.split(':')[-1] for l in labels]
        if labels_map is not None:
            labels = [labels_map.get(l, l) for l in labels]
        if only_markers is not None:
            good = [l in only_markers for l in labels]
        else:
            good = [True] * len(labels)
            if ignore_stared_labels:
                good = [g and not l.startswith('*') for g, l in zip(good, labels)]
            if exclude_markers is not None:
                good = [g and l not in exclude_markers for g, l in zip(good, labels)]
        good = np.asarray(good, dtype=bool)
        labels = [l for l, g in zip(labels, good) if g]
        subject_mask = {k: v[good] for k, v in rd['subject_mask'].items()}
        subject_names = sorted(subject_mask.keys())
        markers = rd['markers'][:, good].copy()
        nan_mask = np.logical_not(MocapSession.marker_availability_mask(markers))
        markers[nan_mask] = 0.
        if mocap_rotate is not None:
            markers = rotate_points_xyz(markers, mocap_rotate).reshape(markers.shape)
        if only_subjects:
            if not np.all([s in subject_names for s in only_subjects]):
                return
            selm = np.zeros(markers.shape[1], dtype=bool)
            for s in only_subjects:
                selm = np.logical_or(selm, subject_mask[s])
            subject_mask = {k: v[selm] for k, v in subject_mask.items() if k in only_subjects}
            subject_names = only_subjects
            markers = markers[:, selm]
            labels = (np.array(labels)[selm]).tolist()
        self.markers = markers / scale
        self.labels = labels
        self.subject_mask = subject_mask
        self.subject_names = subject_names
        self.multi_subject = len([s for s in subject_names if s != 'null']) > 1
        fr = rd.get('frame_rate', 120.)
        self.frame_rate = 120. if fr is None else fr
        self.read_status = True

    @staticmethod
    def marker_availability_mask(markers):
        """A marker is missing if any coordinate is NaN or all three are exactly 0 (line 277)."""
        return np.logical_and(np.isnan(markers).sum(-1) == 0, (markers == 0).sum(-1) != 3)

    def markers_asdict(self) -> List[Dict[str, np.ndarray]]:
        ok = MocapSession.marker_availability_mask(self.markers)
        out = []
        for t in range(self.markers.shape[0]):
            m = OrderedDict()
            for i, l in enumerate(self.labels):
                if ok[t, i]:
                    m[l] = self.markers[t, i, :]
            out.append(m)
        return out

    def frames_for_labels(self, latent_labels: Sequence[str], frame_ids: Sequence[int]):
        """Dense view of what chmosh.py:582-594 builds frame by frame: observations F x M x 3 in
        ``latent_labels`` order and the F x M visibility mask (label present and sample available)."""
        lab_idx = {}
        for i, l in enumerate(self.labels):
            lab_idx[l] = i
        cols = np.array([lab_idx.get(l, -1) for l in latent_labels], dtype=np.int64)
        frame_ids = np.asarray(list(frame_ids), dtype=np.int64)
        mk = self.markers[frame_ids]
        ok = MocapSession.marker_availability_mask(mk)
        have = cols >= 0
        obs = np.zeros((len(frame_ids), len(cols), 3))
        vis = np.zeros((len(frame_ids), len(cols)), dtype=bool)
        obs[:, have] = mk[:, cols[have]]
        vis[:, have] = ok[:, cols[have]]
        obs[~vis] = 0.0
        return obs, vis

    def __len__(self):
        return self.markers.shape[0]

    def __getitem__(self, given):
        return self.markers[given]

    def time_length(self):
        assert self.frame_rate is not None, ValueError(f'mocap frame_rate is unknown: {self.mocap_fname}')
        return self.markers.shape[0] / self.frame_rate
