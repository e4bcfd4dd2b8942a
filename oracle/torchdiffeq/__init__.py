"""TEST INFRASTRUCTURE ONLY -- clean-room CPU restatement of `torchdiffeq` (0.2.x).

The reference (Cosmo-Pop/flowfusion, `pyproject.toml:12-13`) pins the un-vendored PyPI
package ``torchdiffeq >=0.2.5,<0.3.0``; every ODE entry point of the reference calls
``torchdiffeq.odeint`` (`diffusion.py:621,631,734,744`, `flow.py:288,299,358,371,781,792,
855,869`, `symplectic.py:237`).  The package is absent from this image and cannot be
downloaded, so this module restates its *published algorithm* (adaptive Dormand-Prince
5(4) with the Hairer initial-step heuristic, the mixed-precision time bookkeeping, the
dense-output interpolant, and the fixed-grid Euler / midpoint / 3-8-rule RK4 drivers) from
the upstream documentation and behaviour.  PARITY UNPINNED: upstream's own tests cannot be
run against it here; it is pinned instead by (i) scipy's independent RK45 (same tableau),
(ii) closed-form ODE solutions and (iii) the property tests in ``tests/test_oracle.py``.

Only ``tests/``, ``__graft_entry__.smoke()``, ``oracle/make_golden.py`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package.  The
product (``flowfusion_b200``) never does.
"""
from ._solver import odeint, odeint_adjoint, last_stats, SolveStats  # noqa: F401

__version__ = "0.2.5+ffb200.restatement"
