"""Mocap input adapter: file -> dense ``(F x M x 3 observations, F x M visibility)`` in latent-label order.

The Stage-II loop of the reference consumes one ``{label: xyz}`` dictionary per frame
(tools/mocap_interface.py:254-273 ``markers_asdict``, stacked at chmosh.py:582-594).  The device wants the same
information as two dense arrays, so this adapter is built around a column table instead of per-frame dictionaries:

* ``load_markers``   one reader per container (.npz / .pkl / .mat / .c3d via ``c3d_io``), registered by extension;
* ``ColumnTable``    what every input column means after the reference's label clean-up -- blanks removed, subject
                     prefix dropped, synonyms resolved through the label map (chmosh.py:466 always passes
                     ``general_labels_map``; shipped here as data/label_synonyms.tsv), ``*``-labels / excluded /
                     non-selected columns dropped (tools/mocap_interface.py:194-215);
* ``MocapSession``   the reference's attribute surface (``markers`` in metres with missing samples zeroed, ``labels``,
                     ``frame_rate``, ``subject_names``, ``multi_subject``, ``time_length()``) plus
                     ``frames_for_labels`` -- the dense view.  A sample is missing when a coordinate is NaN or all
                     three are exactly zero (:277); of several columns with one label the LAST AVAILABLE one of a
                     frame wins, which is what writing the per-frame dictionary in column order does (:262-271).
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

UNIT_PER_METRE = {'mm': 1000.0, 'cm': 100.0, 'm': 1.0}
_SYNONYM_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'label_synonyms.tsv')
_synonyms: Optional[Dict[str, str]] = None


def general_labels_map() -> Dict[str, str]:
    """raw label -> canonical label (the reference's marker_layout/labels_map.py:34-231, stored grouped by canonical
    label in data/label_synonyms.tsv)."""
    global _synonyms
    if _synonyms is None:
        table: Dict[str, str] = {}
        with open(_SYNONYM_FILE) as f:
            for line in f:
                if line.startswith('#') or not line.strip():
                    continue
                canon, raws = line.rstrip('\n').split('\t')
                for raw in raws.split(' '):
                    table[raw] = canon
        _synonyms = table
    return _synonyms


def rotation_xyz(rxyz_deg: Sequence[float]) -> np.ndarray:
    """Rz Ry Rx for angles in degrees (human_body_prior's ``rotate_points_xyz`` convention)."""
    ax, ay, az = np.radians(np.asarray(rxyz_deg, dtype=np.float64).ravel()[:3])
    rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    return rz @ ry @ rx


def rotate_points_xyz(points: np.ndarray, rxyz_deg: Sequence[float]) -> np.ndarray:
    """F x N x 3 points rotated by Rz Ry Rx (degrees), as human_body_prior's helper of the same name."""
    return np.einsum('cd,fnd->fnc', rotation_xyz(rxyz_deg), points)


# --------------------------------------------------------------------------------------
# containers
# --------------------------------------------------------------------------------------
def _from_npz(fname):
    z = np.load(fname, allow_pickle=True)
    rate = None
    if 'frame_rate' in z.files:
        rate = float(z['frame_rate'])
    elif 'required_parameters' in z.files:
        rate = z['required_parameters'].item().get('frame_rate')
    return z['markers'], (z['labels'].tolist() if 'labels' in z.files else None), rate


def _from_pkl(fname):
    with open(fname, 'rb') as f:
        md = pickle.load(f, encoding='latin-1')
    rate = md['required_parameters']['frame_rate'] if 'required_parameters' in md else md.get('frame_rate')
    labels = md.get('labels', None)
    if isinstance(labels, np.ndarray):
        labels = labels.tolist()
    if labels:   # unlabeled trajectories are stored as arrays: they become starred placeholders
        labels = [f'*{i}' if isinstance(l, np.ndarray) else l for i, l in enumerate(labels)]
    return md['markers'], labels or None, rate


def _from_mat(fname):
    import scipy.io
    md = scipy.io.loadmat(fname)
    keys = [k for k in ('MoCaps', 'Markers') if k in md]
    if not keys:
        raise ValueError("The .mat file do not have the expected field for marker data! "
                         "Expected fields are ['MoCaps', 'Markers']")
    labels = np.vstack(md['Labels'][0]).ravel().tolist() if 'Labels' in md else None
    return md[keys[-1]], labels, None


def _from_c3d(fname):
    from .c3d_io import read_c3d
    return read_c3d(fname)


READERS: Dict[str, Callable] = {'.npz': _from_npz, '.pkl': _from_pkl, '.mat': _from_mat, '.c3d': _from_c3d}


def load_markers(mocap_fname: str) -> Tuple[np.ndarray, List[str], Optional[float]]:
    """(markers F x L x 3 float64 in file units, one label per column, frame rate or None).
    Columns without a label get the placeholder ``*<index>`` (tools/mocap_interface.py:118-124,150-152)."""
    fname = str(mocap_fname)
    ext = os.path.splitext(fname)[1].lower()
    if ext not in READERS:
        raise ValueError(f"Error! Could not recognize file format for {fname}")
    markers, labels, rate = READERS[ext](fname)
    markers = np.array(markers, dtype=np.float64)
    labels = [] if labels is None else [l.decode() if isinstance(l, bytes) else str(l) for l in labels]
    labels += [f'*{i}' for i in range(len(labels), markers.shape[1])]
    return markers, labels, rate


def read_mocap(mocap_fname: str) -> Dict:
    """Dictionary form of ``load_markers`` with the per-subject column masks (tools/mocap_interface.py:87-162)."""
    markers, labels, rate = load_markers(mocap_fname)
    return {'markers': markers, 'labels': labels, 'frame_rate': rate, 'subject_mask': subject_masks(labels)}


def subject_masks(raw_labels: Sequence[str]) -> Dict[str, np.ndarray]:
    """Column mask per capture subject: the text before ':' of a raw label, 'null' without one."""
    names = np.array([l.split(':')[0] if ':' in l else 'null' for l in raw_labels], dtype=object)
    order = list(dict.fromkeys(names.tolist()))
    return {n: names == n for n in order}


# --------------------------------------------------------------------------------------
# label clean-up
# --------------------------------------------------------------------------------------
@dataclass
class ColumnTable:
    labels: List[str]          # cleaned label of every KEPT column
    keep: np.ndarray           # bool mask over the input columns

    @classmethod
    def build(cls, raw_labels: Sequence[str], *, labels_map: Optional[Dict[str, str]], only_markers, exclude_markers,
              ignore_stared_labels: bool, remove_label_before_colon: bool) -> 'ColumnTable':
        clean = [l.replace(' ', '') for l in raw_labels]
        if remove_label_before_colon:
            clean = [l.rsplit(':', 1)[-1] for l in clean]
        if labels_map is not None:
            clean = [labels_map.get(l, l) for l in clean]
        if only_markers is not None:                     # a positive list overrides the two negative rules
            wanted = set(only_markers)
            keep = np.array([l in wanted for l in clean], dtype=bool)
        else:
            banned = set(exclude_markers or ())
            keep = np.array([not (ignore_stared_labels and l.startswith('*')) and l not in banned for l in clean], dtype=bool)
        return cls([l for l, k in zip(clean, keep) if k], keep)


class MocapSession:
    """One capture file, cleaned up like the reference's ``MocapSession`` (tools/mocap_interface.py:165-252).

    ``labels_map``: ``'general'`` (default) = the shipped synonym table the reference's Stage I / II always apply
    (chmosh.py:124,466); a dict = custom table; ``None`` = raw labels."""

    def __init__(self, mocap_fname, mocap_unit: str, mocap_rotate=None, exclude_markers: List[str] = None,
                 only_subjects: List[str] = None, only_markers: List[str] = None, labels_map='general',
                 ignore_stared_labels: bool = True, remove_label_before_colon: bool = True):
        if mocap_unit not in UNIT_PER_METRE:
            raise KeyError(mocap_unit)
        if only_subjects:
            assert isinstance(only_subjects, list), ValueError(
                f'attribute only_subjects should be a list of strings as subject names: {only_subjects}')
        self.mocap_fname = mocap_fname
        self.read_status = False
        raw, raw_labels, rate = load_markers(mocap_fname)
        if isinstance(labels_map, str):
            if labels_map != 'general':
                raise ValueError(f'unknown labels_map {labels_map!r}')
            labels_map = general_labels_map()
        table = ColumnTable.build(raw_labels, labels_map=labels_map, only_markers=only_markers,
                                  exclude_markers=exclude_markers, ignore_stared_labels=ignore_stared_labels,
                                  remove_label_before_colon=remove_label_before_colon)
        subjects = {n: m[table.keep] for n, m in subject_masks(raw_labels).items()}
        labels = table.labels
        raw_cols = np.nonzero(table.keep)[0]                # file column of every kept column
        if only_subjects:
            missing = [s for s in only_subjects if s not in subjects]
            if missing:      # the reference logs an error and leaves the session unread (read_status False)
                return
            cols = np.logical_or.reduce([subjects[s] for s in only_subjects])
            subjects = {s: subjects[s][cols] for s in only_subjects}
            raw_cols = raw_cols[cols]
            labels = [l for l, c in zip(labels, cols) if c]
        # Everything above works on labels only.  The marker array itself -- missing samples zeroed, rotated, in metres --
        # is put together on first use of ``markers`` (the device path uploads the raw table and does this on the GPU,
        # so the clean-up of the host copy can run behind the solve).
        self.raw = raw
        self.raw_columns = raw_cols
        self.unit_per_metre = UNIT_PER_METRE[mocap_unit]
        self.mocap_rotate = mocap_rotate
        self._markers: Optional[np.ndarray] = None
        self.labels = labels
        self.subject_mask = subjects
        self.subject_names = list(only_subjects) if only_subjects else sorted(subjects)
        self.multi_subject = sum(1 for s in self.subject_names if s != 'null') > 1
        self.frame_rate = 120. if rate is None else rate
        self.read_status = True

    @property
    def markers(self) -> np.ndarray:
        """F x L x 3, metres, missing samples zeroed (tools/mocap_interface.py:223-233)."""
        if self._markers is None:
            markers = self.raw[:, self.raw_columns]
            markers[~self.marker_availability_mask(markers)] = 0.0
            if self.mocap_rotate is not None:
                markers = rotate_points_xyz(markers, self.mocap_rotate).reshape(markers.shape)
            self._markers = markers / self.unit_per_metre
        return self._markers

    @markers.setter
    def markers(self, value: np.ndarray):
        self._markers = value

    def raw_columns_for_labels(self, latent_labels: Sequence[str]):
        """File column of every latent label (-1: the file has no such label), for the device-side adapter
        (``Job.upload_markers``).  Returns None when a label owns several columns (the per-frame "last available one wins"
        rule is then applied on the host, ``frames_for_labels``)."""
        columns: Dict[str, List[int]] = {}
        for c, l in enumerate(self.labels):
            columns.setdefault(l, []).append(c)
        out = np.full(len(latent_labels), -1, dtype=np.int32)
        for m, l in enumerate(latent_labels):
            cols = columns.get(l, ())
            if len(cols) > 1:
                return None
            if cols:
                out[m] = self.raw_columns[cols[0]]
        return out

    @staticmethod
    def marker_availability_mask(markers: np.ndarray) -> np.ndarray:
        """False where a sample is missing: a NaN coordinate, or all three exactly 0 (:277)."""
        return ~np.isnan(markers).any(-1) & ~(markers == 0).all(-1)

    def frames_for_labels(self, latent_labels: Sequence[str], frame_ids: Sequence[int]):
        """Observations ``F x M x 3`` (metres) in ``latent_labels`` order and the ``F x M`` visibility mask -- label
        present in the file and sample available in the frame.  What chmosh.py:582-594 stacks frame by frame."""
        if isinstance(frame_ids, range) and len(frame_ids):       # a view, no copy
            mk = self.markers[frame_ids.start:frame_ids.stop:frame_ids.step]
        else:
            mk = self.markers[np.asarray(list(frame_ids), dtype=np.int64)]
        ok = (mk != 0).any(-1)        # missing samples were zeroed at load time (NaN included), so this is the :277 rule
        F, M = mk.shape[0], len(latent_labels)
        obs = np.zeros((F, M, 3))
        vis = np.zeros((F, M), dtype=bool)
        columns: Dict[str, List[int]] = {}
        for c, l in enumerate(self.labels):
            columns.setdefault(l, []).append(c)
        # one gather for the labels that own exactly one column (the normal case) ...
        single = [(m, columns[l][0]) for m, l in enumerate(latent_labels) if len(columns.get(l, ())) == 1]
        if single:
            ms, cs = (np.array(x) for x in zip(*single))
            vis[:, ms] = ok[:, cs]
            obs[:, ms] = np.where(ok[:, cs, None], mk[:, cs], 0.0)
        # ... duplicates: later columns overwrite earlier ones where they are available
        for m, l in enumerate(latent_labels):
            cols = columns.get(l, ())
            if len(cols) > 1:
                for c in cols:
                    sel = ok[:, c]
                    obs[sel, m] = mk[sel, c]
                    vis[:, m] |= sel
        return obs, vis

    def markers_asdict(self) -> List[Dict[str, np.ndarray]]:
        """Per-frame ``{label: xyz}`` of the available samples (compatibility view; the solver uses the dense one)."""
        ok = self.marker_availability_mask(self.markers)
        return [{l: self.markers[t, c] for c, l in enumerate(self.labels) if ok[t, c]} for t in range(len(self))]

    def __len__(self):
        return (self.raw if self._markers is None else self._markers).shape[0]

    def __getitem__(self, given):
        return self.markers[given]

    def time_length(self):
        assert self.frame_rate is not None, ValueError(f'mocap frame_rate is unknown: {self.mocap_fname}')
        return len(self) / self.frame_rate
