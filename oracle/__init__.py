"""CPU float64 restatement of the MoSh++ Stage-II hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package; the product (``moshpp_b200``) never does.

PARITY UNPINNED (except rigid.py, see below).  The reference (nghorbani/moshpp @ 6599a2d) ships no tests, golden vectors or
fixtures for this path, and its arithmetic lives in third-party modules that
are neither vendored nor installable here:

  * ``chumpy`` (requirements.txt:2, unpinned; PyPI latest 0.70) -- ``ch.minimize(method='dogleg')``
    call sites chmosh.py:651-653,669-671,703-705.  Restated in ``dogleg.py`` from the published
    algorithm (chumpy/optimization_internal.py ``_minimize_dogleg`` / ``DoglegState``).
  * ``psbody.smpl`` (MPI-internal, never published) -- ``verts_decorated`` and the C++
    ``lbs_derivatives_wrt_pose/_shape`` called at models/smpl_fast_derivatives.py:206-218,246-263.
    Restated in ``lbs.py`` from the public SMPL formulation (``lrotmin`` pose features, LBS).
  * ``cv2.Rodrigues`` (rigid_transformations.py:82) -- present here; used as an independent check
    of ``rigid.py`` in tests/test_oracle_math.py.

What pins the oracle instead: analytic Jacobians == torch.autograd Jacobians of an
independently written float64 forward; cv2.Rodrigues value + Jacobian; scipy least_squares optimum
cross-check; ground-truth recovery on noise-free synthetic data; committed golden vectors emitted
by this oracle (tests/golden/, generator script tests/golden/make_golden.py).

Pinned by the reference itself: ``moshpp/rigid_transformations.py`` needs only numpy / scipy / cv2 and imports here;
``rigid.py`` is checked against vectors produced by that unmodified module (tests/golden/ref_rigid.npz, generator
tests/golden/make_reference_vectors.py, test tests/test_reference_vectors.py).

Every function cites the reference file:line it follows (paths relative to
/root/reference/src/moshpp unless noted).
"""
