"""latticeqmc_b200 - B200-native engine for LatticeQMC's HS-field Metropolis sweep, behind the
reference's `lqmc` Python API (lqmc/__init__.py:9-96).

The compute path is `liblqmc_b200.so` (hand-written sm_100a CUDA behind the C ABI in
`include/lqmc_b200.h`); importing this package needs neither CUDA nor the library, running a sweep
needs both - there is no CPU fallback.
"""
import numpy as np

from .lattice import Lattice
from .configuration import Configuration
from .hubbard import HubbardModel
from .lqmc import LatticeQMC
from .tools import *  # noqa: F401,F403
from .tools import check_params
from .multiprocessing import ParallelProcessManager, SerialProcessManager, LqmcProcess, ProcessManager
from .engine import SweepEngine, EngineError, philox_uniforms

__version__ = "0.1.0"


def measure(model, beta, time_steps, warmup, sweeps, cores=None, det_mode=False, **engine_kwargs):
    """Warm-up + measurement for one model / temperature; returns `(gf_up, gf_dn)`
    (lqmc/__init__.py:17-54).  `cores` = number of independent chains."""
    if cores is not None and cores == 1:
        solver = LatticeQMC(model, beta, time_steps, warmup, sweeps, det_mode, **engine_kwargs)
        gf_up, gf_dn = solver.run()
    else:
        check_params(model.u, model.t, beta / time_steps)
        manager = ParallelProcessManager(model, beta, time_steps, warmup, det_mode=det_mode, procs=cores, **engine_kwargs)
        manager.set_jobs(sweeps)
        manager.run()
        gf_up, gf_dn = manager.get_result()
    return gf_up, gf_dn


def measure_betas(model, betas, time_steps, warmup=500, sweeps=5000, cores=-1, caching=True, det_mode=False,
                  **engine_kwargs):
    """Beta scan; returns `(gf_up, gf_dn)` each `(M, N, N)` (lqmc/__init__.py:57-96)."""
    manager = SerialProcessManager(model, time_steps, warmup, sweeps, det_mode, cores, caching, **engine_kwargs)
    manager.set_jobs(betas)
    manager.run(sleep=1.0)
    gf_data = manager.get_result()
    gf_up, gf_dn = np.swapaxes(gf_data, 0, 1)
    manager.terminate()
    return gf_up, gf_dn
