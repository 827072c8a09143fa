"""The ED oracle reproduces the reference points measured in the survey: 2x2 (a 4-site ring
with t=1), U=4, beta=2."""
import numpy as np

from oracle import ed
from oracle import sweep_oracle as so


def _ring4():
    ham = so.ideal_square_kinetic(2, 1.0, 0.0)
    np.fill_diagonal(ham, 0.0)
    return ham


def test_ed_half_filling_and_mu4():
    half = ed.thermal_observables(_ring4(), 4.0, 2.0, 2.0)
    assert abs(half["n_up"] - 0.5) < 1e-10 and abs(half["n_dn"] - 0.5) < 1e-10
    assert abs(half["docc"] - 0.08732) < 2e-5 and abs(half["moment"] - 0.82537) < 2e-5
    mu4 = ed.thermal_observables(_ring4(), 4.0, 4.0, 2.0)
    assert abs(mu4["n_up"] - 0.68019) < 2e-5 and abs(mu4["docc"] - 0.38738) < 2e-5


def test_physics_oracle_sweep_is_a_correct_sampler_smoke():
    """One short run of the physics-mode oracle stays in the physical range (a full statistical comparison
    with ED runs on the GPU, tests/test_gpu_physics.py)."""
    from scipy.linalg import expm
    ham = so.ideal_square_kinetic(2, 1.0, 0.0)
    dtau, lamb, exp_k = so.set_beta_constants(ham, 4.0, 2.0, 20)
    exp_k_inv = expm(dtau * ham)
    h = so.initial_field(4, 20, 1)
    rs = np.random.RandomState(2)
    obs = {}
    for _ in range(30):
        so.physics_sweep(h, exp_k, exp_k_inv, lamb, rs.rand(20, 4), observables=obs)
    n = (obs["n_up"] + obs["n_dn"]) / obs["count"]
    assert 0.7 < n < 1.3 and 0.0 < obs["docc"] / obs["count"] < 0.3
