"""Device-resident NVE molecular dynamics on the continuation surface, batched over
independent trajectories.

The reference runs ``pyscf.md.NVE`` on the host with one ``get_energy_with_grad`` call per
step (evcont/MD_utils.py:60-125).  Here the whole step -- positions update, AO integrals
(K9), prediction (K3..K8), velocities update, frame recording -- is enqueued on one CUDA
stream with no host synchronisation, optionally captured once in a CUDA graph and replayed.

Velocity Verlet as ``pyscf.md.integrators.VelocityVerlet`` does it (PySCF is not in this
image, so this is pinned against the numpy restatement in ``oracle/md.py``, not against
PySCF itself): frame 0 is the initial geometry; ``x += dt v + dt^2/2 a``;
``a' = -grad/m``; ``v += dt/2 (a + a')``.  Masses default to the most common isotope
(``pyscf.data.elements.COMMON_ISOTOPE_MASSES``) in atomic units.
"""
import numpy as np
import torch

from .stackcache import as_device_stack

AMU2AU = 1822.888486209  # pyscf.data.nist.AMU2AU
COMMON_ISOTOPE_MASSES = {"H": 1.00782503223, "He": 4.00260325413, "O": 15.99491461957}


def atomic_masses(mol):
    return np.array([COMMON_ISOTOPE_MASSES[mol.atom_symbol(i)] for i in range(mol.natm)]) * AMU2AU


class DeviceNVE:
    """``B`` independent NVE trajectories of one molecule, advanced together on one GPU.

    ``mol``: :class:`evcont_b200.mol.MolLite` (atoms + basis); ``coords0``: ``(B, natm, 3)`` bohr;
    ``veloc0``: same shape or ``None`` (zero); the stack as for ``get_energy_with_grad``.
    """

    def __init__(self, mol, one_rdm, two_rdm, overlap, coords0, veloc0=None, dt=10.0, masses=None,
                 max_frames=0, use_graph=True):
        self.stack = as_device_stack(one_rdm, two_rdm, overlap)
        eng = self.engine = self.stack.engine
        self.mol, self.dt = mol, float(dt)
        self.sbasis = mol.sbasis(eng)
        coords0 = np.ascontiguousarray(coords0, dtype=np.float64).reshape(-1, mol.natm, 3)
        B, natm, N = coords0.shape[0], mol.natm, self.stack.ntrain
        self.nbatch = B
        m = atomic_masses(mol) if masses is None else np.asarray(masses, dtype=np.float64)
        self.mass, self.inv_mass = eng.to_device(m), eng.to_device(1.0 / m)
        self.x = eng.to_device(coords0)
        self.v = eng.to_device(np.zeros_like(coords0) if veloc0 is None
                               else np.ascontiguousarray(veloc0, dtype=np.float64).reshape(B, natm, 3))
        self.a = torch.zeros_like(self.x)
        self.epot, self.ekin = eng.empty(B), eng.empty(B)
        self.grad, self.cvec = eng.empty(B, natm, 3), eng.empty(B, N)
        from .engine import DeviceAO
        self.ao = DeviceAO(eng, B, self.sbasis.nao, natm, self.sbasis.aoslices_host,
                           packed=(self.sbasis.nao <= 13 and not self.sbasis.general))
        self.max_frames = int(max_frames)
        self.frame_idx = torch.zeros(1, dtype=torch.int32, device=eng.device)
        mf = max(1, self.max_frames)
        self.traj = eng.empty(mf, B, natm, 3) if self.max_frames else None
        self.epot_log = eng.empty(mf, B) if self.max_frames else None
        self.ekin_log = eng.empty(mf, B) if self.max_frames else None
        self.use_graph, self._graph = bool(use_graph), None
        self.nsteps = 0
        # scratch owned by this object: the captured graph replays with this buffer's address, so it must
        # neither be freed nor shared with other engine calls issued between two run() calls
        from .engine import Workspace
        self._ws = Workspace(eng.device)
        self._force(first=True)

    def _force(self, first):
        eng = self.engine
        with eng.using_workspace(self._ws):
            self._force_impl(first)

    def _force_impl(self, first):
        eng = self.engine
        eng.energy_with_grad_coords(self.stack, self.sbasis, self.x, ao=self.ao,
                                    out=(self.epot, self.grad, self.cvec))
        eng.md_velocities(self.dt, first, self.inv_mass, self.mass, self.grad, self.x, self.epot, self.v,
                          self.a, self.ekin, self.frame_idx, self.max_frames, self.traj, self.epot_log,
                          self.ekin_log)

    def _thermostat(self):
        """Hook called before the positions update (NVE: nothing)."""

    def _step(self):
        self._thermostat()
        self.engine.md_positions(self.dt, self.x, self.v, self.a)
        self._force(first=False)

    def run(self, steps):
        """Advance every trajectory by ``steps`` steps (asynchronous; results are read with
        :meth:`frames` / ``.x`` after a stream synchronisation, which those accessors do)."""
        if steps <= 0:
            return self
        if self.use_graph and self._graph is None:
            self._step()  # eager once: workspaces allocated, attributes set
            steps -= 1
            self.nsteps += 1
            torch.cuda.synchronize(self.engine.device)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(self.engine.device)
            side.wait_stream(torch.cuda.current_stream(self.engine.device))
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    self._step()
            torch.cuda.current_stream(self.engine.device).wait_stream(side)
            # capturing does not execute: the captured step has not been applied
            self._graph = g
            self._ws.freeze()
        for _ in range(steps):
            if self._graph is not None:
                self._graph.replay()
            else:
                self._step()
        self.nsteps += steps
        return self

    def frames(self):
        """``(trajectory[frames, B, natm, 3], epot[frames, B], ekin[frames, B])`` as numpy."""
        torch.cuda.synchronize(self.engine.device)
        n = min(int(self.frame_idx.item()), self.max_frames)
        return (self.traj[:n].cpu().numpy(), self.epot_log[:n].cpu().numpy(), self.ekin_log[:n].cpu().numpy())


class DeviceNVT(DeviceNVE):
    """Berendsen-thermostatted trajectories (``pyscf.md.NVTBerendson(scanner, T=, taut=)``, as in the
    reference's Zundel scripts): the velocities are rescaled by
    ``clip(sqrt(1 + (T/T_inst - 1) dt/taut), 0.9, 1.1)`` before every step."""

    def __init__(self, mol, one_rdm, two_rdm, overlap, coords0, veloc0=None, dt=10.0, T=298.15, taut=250.0,
                 **kwargs):
        self.T, self.taut = float(T), float(taut)
        super().__init__(mol, one_rdm, two_rdm, overlap, coords0, veloc0, dt=dt, **kwargs)

    def _thermostat(self):
        self.engine.md_berendsen(self.dt, self.taut, self.T, self.ekin, self.v)
