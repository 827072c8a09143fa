"""Hubbard model on a finite lattice: produces the dense kinetic matrix ``K`` the sweep consumes.

Host-side input producer for the hot path; API mirrors the reference ``lqmc/hubbard.py:12-111``
(``HubbardModel(u, t, mu)``, ``set_params``, ``param_str``, ``n_sites``, ``build``, ``build_square``,
``ham_kinetic``).  Convention kept from the reference (``hubbard.py:30,99-100``): ``mu`` defaults to
``u/2`` and the diagonal of ``K`` is ``-mu``; together with the reference's HS decoupling this
simulates ``mu_true = mu + u/2``.  Nothing here or in the engine corrects that silently: true
half filling in physics mode is requested by building the model with ``mu=0`` (``tests/test_gpu_physics.py``).
"""
import numpy as np

from .lattice import Lattice


class HubbardModel:

    def __init__(self, u=4, t=1, mu=None):
        self.lattice = Lattice.square()
        self.u = u
        self.t = t
        self.mu = u / 2 if mu is None else mu

    def set_params(self, u, t=1.0, mu=None):
        self.u = u
        self.t = t
        self.mu = u / 2 if mu is None else mu

    def param_str(self):
        return f"u={self.u}_t={self.t}_mu={self.mu}"

    def __str__(self):
        return f"HubbardModel(u={self.u}, t={self.t}, mu={self.mu})"

    @property
    def n_sites(self):
        return self.lattice.n_sites

    def build(self, width, height=1, cycling=0):
        """Rectangular ``width x height`` lattice; ``cycling`` = axis/axes closed periodically
        (``None`` for open boundaries).  Reference: ``hubbard.py:60-76``."""
        self.lattice.build((width, height))
        if cycling is not None:
            self.lattice.set_periodic_boundary(cycling)

    def build_square(self, size, cycling=(0, 1)):
        self.build(size, size, cycling)

    def ham_kinetic(self):
        """Dense ``(N, N)`` float64 one-particle matrix: ``-mu`` on the diagonal, ``-t`` on every
        nearest-neighbour bond (assigned, not accumulated: a doubly-listed bond of a 2-site ring
        still carries a single ``-t``, as in the reference ``hubbard.py:91-111``)."""
        n = self.lattice.n_sites
        ham = -self.mu * np.eye(n, dtype=np.float64)
        for i in range(n):
            for j in self.lattice.nearest_neighbours(i):
                if j < i:
                    ham[i, j] = -self.t
                    ham[j, i] = -np.conj(self.t)
        return ham
