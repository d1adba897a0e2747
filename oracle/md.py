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
COMMON_ISOTOPE_MASSES = {"H": 1.00782503223, "He": 4.00260325413, "O": 15.99491461957}


KB_HARTREE = 3.166811563e-6


def velocity_verlet(coords0, veloc0, masses, dt, steps, energy_grad, berendsen=None):
    """``energy_grad(coords) -> (E, grad)``; returns ``(traj[steps], epot[steps], ekin[steps])``.
    ``berendsen=(T, taut)`` rescales the velocities before every step as pyscf.md.NVTBerendson does:
    ``v *= clip(sqrt(1 + (T/T_inst - 1) dt/taut), 0.9, 1.1)``, ``T_inst = 2 E_kin / (3 natm k_B)``."""
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
            if berendsen is not None:
                tinst = 2.0 * 0.5 * (m * v * v).sum() / (x.size * KB_HARTREE)
                f = 1.1 if tinst <= 0 else float(np.clip(np.sqrt(max(1.0 + (berendsen[0] / tinst - 1.0) * dt
                                                                   / berendsen[1], 0.81)), 0.9, 1.1))
                v = v * f
            x = x + dt * v + 0.5 * dt * dt * a
            e, g = energy_grad(x)
            an = -g / m
            v = v + 0.5 * dt * (a + an)
            a = an
        traj.append(x.copy())
        epot.append(e)
        ekin.append(0.5 * (m * v * v).sum())
    return np.array(traj), np.array(epot), np.array(ekin)
