"""Velocity-Verlet NVE integration -- CPU oracle for evcont_b200.md (TEST INFRASTRUCTURE ONLY).

Restates ``pyscf.md.integrators.VelocityVerlet`` as the reference drives it through
``md.NVE(scanner, dt, steps, veloc, frames=frames).run()`` (evcont/MD_utils.py:104-119):
frame 0 holds the initial geometry (only the acceleration is evaluated), then
``x += dt v + dt^2/2 a``, ``a' = -grad(x)/m``, ``v += dt/2 (a + a')``.  PySCF is absent from
this image: **parity unpinned with respect to pyscf.md**; the restatement is checked by energy
conservation and time reversibility (tests/test_gpu_md.py).
"""
import numpy as np

AMU2AU = 1822.888486209
COMMON_ISOTOPE_MASSES = {"H": 1.00782503223, "He": 4.00260325413}


def velocity_verlet(coords0, veloc0, masses, dt, steps, energy_grad):
    """``energy_grad(coords) -> (E, grad)``; returns ``(traj[steps], epot[steps], ekin[steps])``."""
    x = np.array(coords0, dtype=np.float64)
    v = np.zeros_like(x) if veloc0 is None else np.array(veloc0, dtype=np.float64)
    m = np.asarray(masses, dtype=np.float64)[:, None]
    traj, epot, ekin = [], [], []
    a = None
    for _ in range(steps):
        if a is None:
            e, g = energy_grad(x)
            a = -g / m
        else:
            x = x + dt * v + 0.5 * dt * dt * a
            e, g = energy_grad(x)
            an = -g / m
            v = v + 0.5 * dt * (a + an)
            a = an
        traj.append(x.copy())
        epot.append(e)
        ekin.append(0.5 * (m * v * v).sum())
    return np.array(traj), np.array(epot), np.array(ekin)
