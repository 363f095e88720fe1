"""TEST-SIDE stand-in for `chumpy`, forward evaluation only -- just enough for the UNMODIFIED reference files
prior/gmm_prior_ch.py and transformed_lm.py to run in this container (chumpy itself is not installable here, SURVEY.md
8(c)) so that they can emit golden vectors (tests/golden/make_reference_vectors.py).  No derivatives, no caching, no
dependency tracking: a `Ch` is a numpy array with chumpy's calling conventions:

  * keyword construction of subclasses: every keyword becomes an attribute, arrays given for names in `dterms` are wrapped,
    then `on_changed(all dterms)` runs; `.r` is `compute_r()` where a subclass defines it, else the wrapped array;
  * arrays stay at least one-dimensional (`x.sum()` has shape (1,), `x[i]` of a vector has shape (1,)), as in chumpy.

Never imported by moshpp_b200 or by the oracle."""
import numpy as np

__all__ = ['Ch', 'array', 'asarray', 'concatenate', 'hstack', 'vstack', 'sqrt', 'sum', 'cross']


def _val(x):
    return x.r if isinstance(x, Ch) else np.asarray(x, dtype=np.float64)


class Ch:
    dterms = ()
    __array_priority__ = 1000.0

    def __init__(self, *args, **kwargs):
        if args:
            object.__setattr__(self, '_x', np.atleast_1d(np.array(_val(args[0]), dtype=np.float64)))
        for k, v in kwargs.items():
            setattr(self, k, v)

    @classmethod
    def _class_dterms(cls):
        d = cls.dterms
        return (d,) if isinstance(d, str) else tuple(d)

    def _names(self):
        return self.__dict__.get('_dterm_names', self._class_dterms())

    def __setattr__(self, name, value):
        # chumpy semantics: assigning a differentiable term wraps arrays and marks the term as changed; `on_changed(which)`
        # runs before the next evaluation
        if name in self._names():
            if not isinstance(value, Ch):
                value = Ch(value)
            self.__dict__.setdefault('_dirty', []).append(name)
        object.__setattr__(self, name, value)

    def add_dterm(self, name, value):
        object.__setattr__(self, '_dterm_names', tuple(self._names()) + (name,))
        object.__setattr__(self, name, value)

    def _flush(self):
        dirty = self.__dict__.get('_dirty')
        if dirty and hasattr(self, 'on_changed'):
            which = list(dict.fromkeys(dirty))
            self.__dict__['_dirty'] = []
            self.on_changed(which)

    # ---- value
    @property
    def r(self):
        self._flush()
        if type(self) is not Ch and hasattr(self, 'compute_r'):
            return np.atleast_1d(np.asarray(_val(self.compute_r()), dtype=np.float64))
        return self._x

    def __array__(self, dtype=None, copy=None):
        return self.r if dtype is None else self.r.astype(dtype)

    @property
    def shape(self):
        return self.r.shape

    def __len__(self):
        return len(self.r)

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __getitem__(self, idx):
        return Ch(np.atleast_1d(self.r[idx]))

    def reshape(self, *shape):
        return Ch(self.r.reshape(*shape))

    def ravel(self):
        return Ch(self.r.ravel())

    def dot(self, other):
        return Ch(self.r.dot(_val(other)))

    def sum(self, axis=None):
        return Ch(np.atleast_1d(self.r.sum(axis=axis)))

    # ---- arithmetic
    def __add__(self, o): return Ch(self.r + _val(o))
    def __radd__(self, o): return Ch(_val(o) + self.r)
    def __sub__(self, o): return Ch(self.r - _val(o))
    def __rsub__(self, o): return Ch(_val(o) - self.r)
    def __mul__(self, o): return Ch(self.r * _val(o))
    def __rmul__(self, o): return Ch(_val(o) * self.r)
    def __truediv__(self, o): return Ch(self.r / _val(o))
    def __rtruediv__(self, o): return Ch(_val(o) / self.r)
    def __pow__(self, p): return Ch(self.r ** p)
    def __neg__(self): return Ch(-self.r)


def array(x):
    return Ch(x)


def asarray(x):
    return x if isinstance(x, Ch) else Ch(np.asarray([_val(v) for v in x]) if isinstance(x, (list, tuple)) else x)


def concatenate(parts, axis=0):
    return Ch(np.concatenate([np.atleast_1d(_val(p)) for p in parts], axis=axis))


def hstack(parts):
    return Ch(np.hstack([_val(p) for p in parts]))


def vstack(parts):
    return Ch(np.vstack([_val(p) for p in parts]))


def sqrt(x):
    with np.errstate(invalid='ignore'):
        return Ch(np.sqrt(_val(x)))


def sum(x, axis=None):
    return Ch(np.atleast_1d(_val(x).sum(axis=axis)))


def cross(a, b):
    return Ch(np.cross(_val(a), _val(b)))
