"""Hubbard-Stratonovich Ising field container.

The layout of ``Configuration.config`` - ``np.int8``, shape ``(n_sites, time_steps)``, C order,
values +-1 - is the host-side input/output contract of the sweep kernels (the engine transposes
to slice-major on upload).  API mirrors the reference ``lqmc/configuration.py:86-188``; the
initial field uses the same draw from the legacy global NumPy stream
(``2*randint(0,2,(N,L))-1``, ``configuration.py:123-124``) so seeded runs start identically.
The matplotlib plot classes of the reference (``configuration.py:12-83``) are out of scope;
``show`` imports matplotlib lazily.
"""
import numpy as np


class Configuration:

    dtype = np.int8

    def __init__(self, n_sites, time_steps, array=None):
        self.n_sites = n_sites
        self.time_steps = time_steps
        self.config = None
        if array is None:
            self.initialize()
        else:
            self.config = array

    def copy(self):
        return Configuration(self.n_sites, self.time_steps, array=self.config.copy())

    def initialize(self):
        bits = np.random.randint(0, 2, size=(self.n_sites, self.time_steps))
        self.config = (2 * bits - 1).astype(self.dtype)

    def update(self, i, t):
        self.config[i, t] *= -1

    def get(self, i, t):
        return self.config[i, t]

    def mean(self):
        return np.mean(self.config)

    def var(self):
        return np.var(self.config)

    def __eq__(self, other):
        return np.all(self.config == other.config)

    def __getitem__(self, item):
        return self.config[item]

    def string_header(self, delim=" "):
        return r"i\l  " + delim.join(f"{l:^3}" for l in range(self.time_steps))

    def string_bulk(self, delim=" "):
        lines = []
        for site in range(self.n_sites):
            cells = delim.join(f"{x:^3}" for x in self.config[site, :])
            lines.append(f"{site:<3} [{cells}]")
        return "\n".join(lines)

    def __str__(self):
        return self.string_header(" ") + "\n" + self.string_bulk(" ")

    def show(self, show=True):  # pragma: no cover - visualisation is out of scope
        import matplotlib.pyplot as plt
        fig, ax = plt.subplots()
        ax.imshow(self.config, cmap="Greys", aspect="auto")
        ax.set_xlabel("time slice")
        ax.set_ylabel("site")
        if show:
            plt.show()
        return fig
