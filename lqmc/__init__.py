"""`import lqmc` drop-in: the reference's package name, served by latticeqmc_b200."""
import sys

import latticeqmc_b200 as _impl
from latticeqmc_b200 import *  # noqa: F401,F403
from latticeqmc_b200 import (Lattice, Configuration, HubbardModel, LatticeQMC, ParallelProcessManager,
                             SerialProcessManager, measure, measure_betas)
from latticeqmc_b200 import lattice, configuration, hubbard, tools, logging, multiprocessing
from latticeqmc_b200 import lqmc as _lqmc_module

for _name, _mod in (("lattice", lattice), ("configuration", configuration), ("hubbard", hubbard), ("tools", tools),
                    ("logging", logging), ("multiprocessing", multiprocessing), ("lqmc", _lqmc_module)):
    sys.modules[f"lqmc.{_name}"] = _mod
lqmc = _lqmc_module
