"""Finite Bravais lattices: geometry -> neighbour lists.

Host-side setup only (runs once per model; its output feeds ``HubbardModel.ham_kinetic``).
Mirrors the public surface of the reference ``lqmc/lattice.py:283-1033`` (class ``Lattice``:
prefabs ``chain/square/hexagonal/cubic``, ``add_atom``, ``calculate_distances``, ``build``,
``set_periodic_boundary``, ``nearest_neighbours``, ``position`` and the cached attributes
``n_sites, shape, indices, neighbours, distances``) but is built differently:

* sites live in an integer-coordinate hash map, so neighbour lookup is O(1) per bond instead of
  the reference's windowed ``np.where`` scan (``lattice.py:771-810``);
* periodic boundaries wrap the *lattice index* modulo the shape.  The reference matches
  wrap-around pairs by Euclidean distance ``== L-1`` (``lattice.py:1027-1031``), which also fires
  for Pythagorean offsets ((9,12) at L=16) and raises on numpy >= 2.2, so it cannot build the
  16x16 lattice at all.  Wherever the reference does build (L = 2,3,4,5,8,24,
  rings) the resulting kinetic matrix is identical; site order is the reference's
  ``index = x*height + y`` (``lattice.py:226-242,886-892``).

Plotting (``lattice.py:16-159,1035-1098``) is out of scope; ``show`` imports matplotlib lazily.
"""
import itertools

import numpy as np


def distance(r1, r2):
    """Euclidean distance between two points."""
    d = np.asarray(r1, dtype=float) - np.asarray(r2, dtype=float)
    return float(np.sqrt(np.dot(d, d)))


def vrange(axis_ranges):
    """All integer vectors of the Cartesian product, first axis slowest (site order contract)."""
    return [np.array(v) for v in itertools.product(*[list(r) for r in axis_ranges])]


class Lattice:

    DIST_DECIMALS = 5

    def __init__(self, vectors):
        self.vectors = np.asarray(vectors, dtype=float)
        self.dim = len(self.vectors)
        # unit cell
        self.distances = []
        self._base_neighbors = []      # [alpha][dist_idx] -> list of int arrays [dn..., alpha']
        self.atoms = []
        self.atom_positions = []
        # finite-size cache
        self.n_sites = 0
        self.shape = None
        self.indices = None
        self.neighbours = None
        self._lookup = {}
        self._periodic = ()

    # ------------------------------------------------------------------ prefabs
    @classmethod
    def chain(cls, name="A", a=1., neighbour_dist=1, size=None):
        latt = cls(np.eye(2) * a)
        latt.add_atom(name=name)
        latt.calculate_distances(neighbour_dist)
        if size is not None:
            latt.build((size, 1))
        return latt

    @classmethod
    def square(cls, name="A", a=1., neighbour_dist=1, shape=None):
        latt = cls(np.eye(2) * a)
        latt.add_atom(name=name)
        latt.calculate_distances(neighbour_dist)
        if shape is not None:
            latt.build(shape)
        return latt

    @classmethod
    def hexagonal(cls, atom1="A", atom2="B", a=1., neighbour_dist=1, shape=None):
        vectors = a * np.array([[np.sqrt(3), np.sqrt(3) / 2], [0, 3 / 2]])
        latt = cls(vectors)
        latt.add_atom(atom1)
        latt.add_atom(atom2, pos=[0, a])
        latt.calculate_distances(neighbour_dist)
        if shape is not None:
            latt.build(shape)
        return latt

    @classmethod
    def cubic(cls, name="A", a=1., neighbour_dist=1, shape=None):
        latt = cls(np.eye(3) * a)
        latt.add_atom(name=name)
        latt.calculate_distances(neighbour_dist)
        if shape is not None:
            latt.build(shape)
        return latt

    # ------------------------------------------------------------------ unit cell
    @property
    def n_base(self):
        return len(self.atoms)

    @property
    def n_dist(self):
        return len(self.distances)

    def __str__(self):
        lines = ["".join(self.atoms) + "-Lattice:"]
        for k, (atom, pos) in enumerate(zip(self.atoms, self.atom_positions)):
            lines.append(f"   {k + 1}: '{atom}' @ {pos}")
        lines.append("   Distances: " + ", ".join(str(d) for d in self.distances))
        return "\n".join(lines) + "\n"

    def add_atom(self, name="A", pos=None, neighbour_dist=0):
        pos = np.zeros(self.dim) if pos is None else np.asarray(pos, dtype=float)
        for other in self.atom_positions:
            if np.array_equal(other, pos):
                raise ValueError(f"Position {pos} allready occupied")
        self.atoms.append(name)
        self.atom_positions.append(pos)
        if neighbour_dist:
            self.calculate_distances(neighbour_dist)

    def get_atom(self, alpha):
        return self.atoms[alpha]

    def get_position(self, n, alpha=0):
        return self.atom_positions[alpha] + self.vectors @ np.asarray(n, dtype=float)

    def calculate_distances(self, n=1):
        """Find the ``n`` shortest inter-site distances and, per basis atom and distance shell,
        the relative lattice indices ``[dn..., alpha']`` of its neighbours."""
        reach = n + 1
        cells = vrange(self.dim * [range(-reach, reach + 1)])
        shells = set()
        for alpha in range(self.n_base):
            r0 = self.get_position(np.zeros(self.dim), alpha)
            for cell in cells:
                for beta in range(self.n_base):
                    d = round(distance(r0, self.get_position(cell, beta)), self.DIST_DECIMALS)
                    if d > 0:
                        shells.add(d)
        self.distances = sorted(shells)[:n]
        self._base_neighbors = []
        for alpha in range(self.n_base):
            r0 = self.get_position(np.zeros(self.dim), alpha)
            per_shell = []
            for dist in self.distances:
                found = []
                for cell in cells:
                    for beta in range(self.n_base):
                        d = round(distance(r0, self.get_position(cell, beta)), self.DIST_DECIMALS)
                        if d == dist:
                            found.append(np.array([*cell, beta], dtype=int))
                per_shell.append(found)
            self._base_neighbors.append(per_shell)

    def get_neighbours(self, idx, dist_idx=0):
        """Lattice indices of the neighbours of site ``idx = [n..., alpha]`` in an infinite lattice."""
        idx = np.asarray(idx, dtype=int)
        out = []
        for rel in self._base_neighbors[int(idx[-1])][dist_idx]:
            nb = rel.copy()
            nb[:-1] += idx[:-1]
            out.append(nb)
        return out

    # ------------------------------------------------------------------ finite lattice cache
    @property
    def n_cells(self):
        return int(np.prod(self.shape))

    @property
    def n(self):
        return len(self.indices) if self.indices is not None else 0

    def lattice_index(self, i):
        idx = self.indices[i]
        return idx[:-1], idx[-1]

    def alpha(self, i):
        return self.indices[i][-1]

    def position(self, i):
        n, alpha = self.lattice_index(i)
        return self.get_position(n, alpha)

    def dist_neighbours(self, i, dist=1):
        return self.neighbours[i][dist - 1]

    def nearest_neighbours(self, i):
        return self.neighbours[i][0]

    def get_list_idx(self, n, alpha=0):
        return self._lookup[(*[int(x) for x in n], int(alpha))]

    def _connect(self, periodic_axes=()):
        """(Re)build the neighbour lists of the cached sites; indices along ``periodic_axes``
        wrap modulo the shape."""
        shape = [int(s) for s in self.shape]
        neighbours = []
        for i_site, idx in enumerate(self.indices):
            per_shell = []
            for i_dist in range(self.n_dist):
                found = []
                for nb in self.get_neighbours(idx, i_dist):
                    key = [int(x) for x in nb]
                    for ax in periodic_axes:
                        key[ax] %= shape[ax]
                    j = self._lookup.get(tuple(key))
                    if j is not None and j != i_site:
                        found.append(j)
                per_shell.append(found)
            neighbours.append(per_shell)
        self.neighbours = neighbours

    def build(self, shape):
        """Cache a finite lattice of ``shape`` cells with open boundaries."""
        shape = np.asarray(shape, dtype=int)
        rows = [[*cell, alpha] for cell in vrange([range(s) for s in shape]) for alpha in range(self.n_base)]
        self.shape = shape
        self.indices = np.array(rows, dtype=int)
        self.n_sites = len(rows)
        self._lookup = {tuple(int(x) for x in row): i for i, row in enumerate(rows)}
        self._periodic = ()
        self._connect()
        return self.indices, shape

    def build_rect(self, width, height):
        return self.build((width, height))

    def add_slices(self, n):
        """Grow the cached lattice by ``n`` cells along the first axis."""
        shape = np.array(self.shape)
        shape[0] += n
        periodic = self._periodic
        self.build(shape)
        if periodic:
            self.set_periodic_boundary(periodic)
        return self.indices, self.neighbours

    def set_periodic_boundary(self, axis=0):
        """Close the cached lattice along one or several axes (index wrap, see module docstring)."""
        axes = tuple(sorted(set(int(a) for a in np.atleast_1d(axis)) | set(self._periodic)))
        self._periodic = axes
        self._connect(axes)

    def show(self, show=True, **kwargs):  # pragma: no cover - visualisation is out of scope
        import matplotlib.pyplot as plt
        fig, ax = plt.subplots()
        pos = np.array([self.position(i) for i in range(self.n_sites)])
        for i in range(self.n_sites):
            for j in self.nearest_neighbours(i):
                if i < j:
                    ax.plot(*zip(pos[i], pos[j]), color="k", lw=kwargs.get("lw", 1.))
        ax.scatter(pos[:, 0], pos[:, 1], zorder=3)
        ax.set_aspect("equal")
        if show:
            plt.show()
        return fig
