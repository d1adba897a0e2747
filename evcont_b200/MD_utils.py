"""``get_scanner``: the drop-in boundary for MD drivers (evcont/MD_utils.py:20-57).

Returns an object PySCF's ``md.NVE`` / ``md.NVTBerendson`` (or any velocity-Verlet
loop) can call as ``scanner(mol) -> (E_total, grad)``.  ``get_trajectory`` runs the NVE
trajectory itself on the device (``evcont_b200.md.DeviceNVE``) when given a
:class:`evcont_b200.mol.MolLite`; ``converge_EVCont_MD`` (the active-learning outer loop,
evcont/MD_utils.py:128-502) keeps the reference's control flow and files and runs each of its
heavy steps batched on the device: the trajectory (``DeviceNVE``), the energies of every frame
with the previous training set (one batched prediction instead of one call per frame), the
farthest-point selection in Hamiltonian space (``evc_min_sqdist``) and the FCI solve + t-RDM
growth of ``FCI_EVCont_obj.append_to_rdms``.
"""
import os

import numpy as np

from .ab_initio_gradients_loewdin import get_energy_with_grad
from .mol import ao_bundle

try:  # subclass PySCF's GradScanner when it exists so isinstance checks in pyscf.md pass
    from pyscf.lib import GradScanner as _ScannerBase
except ImportError:
    _ScannerBase = object


def get_scanner(mol, one_rdm, two_rdm, overlap, hermitian=True):
    """Fake gradient scanner over an eigenvector continuation (same attributes as the
    reference: ``.mol``, ``.base.converged/ovlp/one_trdm/two_trdm/predicted_*``)."""

    class Base:
        converged = True
        ovlp = overlap
        one_trdm = one_rdm
        two_trdm = two_rdm
        predicted_one_rdm = None
        predicted_two_rdm = None

    class Scanner(_ScannerBase):
        def __init__(self):
            self.mol = mol
            self.base = Base()

        def __call__(self, mol):
            self.mol = mol
            if one_rdm is not None and two_rdm is not None and overlap is not None:
                en, grad, rdm_o, rdm_t = get_energy_with_grad(
                    mol, one_rdm, two_rdm, overlap, hermitian=hermitian,
                    return_density_matrices=True)
                self.base.predicted_one_rdm = rdm_o
                self.base.predicted_two_rdm = rdm_t
                return en, grad
            b = ao_bundle(mol)
            return b["e_nuc"], b["grad_nuc"]

    return Scanner()


def write_xyz_trajectory(out, symbols, traj, times):
    """Frames in the xyz form ``pyscf.md`` integrators write to ``trajectory_output`` (``_write_coord``: the
    atom count, ``MD Time <t>``, then ``mol.tostring(format="raw")``: symbol and Cartesian coordinates in
    Angstrom).  ``traj`` (nframes, natm, 3) in bohr, ``times`` in atomic units.  [Format from recollection of
    pyscf/md/integrators.py -- PySCF is not installed here; any xyz reader parses it.]"""
    from .mol import BOHR
    close = isinstance(out, (str, os.PathLike))
    fl = open(out, "w") if close else out
    try:
        for frame, t in zip(np.asarray(traj) * BOHR, times):
            fl.write("%s\nMD Time %.2f\n" % (len(symbols), t))
            for sym, (x, y, z) in zip(symbols, frame):
                fl.write("%-4s %17.8f %17.8f %17.8f\n" % (sym, x, y, z))
        fl.flush()
    finally:
        if close:
            fl.close()


def write_md_energies(out, times, epot, ekin):
    """The table ``pyscf.md`` integrators write to ``energy_output`` / ``data_output`` (``_write_energy``): a
    header line, then ``time  Epot  Ekin  Etot`` per step as ``%8.2f  %.12E  %.12E  %.12E`` (atomic units).
    [Format from recollection of pyscf/md/integrators.py.]  Read it back with :func:`read_md_energies`."""
    close = isinstance(out, (str, os.PathLike))
    fl = open(out, "w") if close else out
    try:
        fl.write("   time          Epot                 Ekin                 Etot\n")
        for t, ep, ek in zip(times, epot, ekin):
            fl.write("%8.2f  %.12E  %.12E  %.12E\n" % (t, ep, ek, ep + ek))
        fl.flush()
    finally:
        if close:
            fl.close()


def read_md_energies(path):
    """(nsteps, 4) table ``time, Epot, Ekin, Etot`` of an energy file, with or without the header line (the
    reference reads column 1 with ``np.genfromtxt``, evcont/MD_utils.py:208)."""
    tab = np.atleast_2d(np.genfromtxt(path))
    return tab[~np.isnan(tab).any(axis=1)]


def get_trajectory(init_mol, overlap, one_rdm, two_rdm, dt=10.0, steps=10, init_veloc=None, hermitian=True,
                   trajectory_output=None, energy_output=None):
    """MD trajectory from the continuation, ``(steps, natm, 3)`` in bohr (evcont/MD_utils.py:60-125:
    ``pyscf.md.NVE`` with ``frames=[]``; frame 0 is the initial geometry).  The integrator runs on
    the device; ``trajectory_output`` / ``energy_output`` (file names or file objects) receive the files
    ``pyscf.md.NVE`` writes: xyz frames headed ``MD Time <t>`` and the ``time Epot Ekin Etot`` table."""
    if hermitian is not True:
        raise NotImplementedError("hermitian=False is not implemented on the device")
    from .md import DeviceNVE
    from .mol import MolLite
    init_mol = MolLite.from_mol(init_mol)  # a PySCF-like Mole with a named basis is re-expressed for the device
    veloc = None if init_veloc is None else np.asarray(init_veloc, dtype=np.float64)[None]
    nve = DeviceNVE(init_mol, one_rdm, two_rdm, overlap, init_mol.atom_coords()[None], veloc, dt=dt,
                    max_frames=steps)
    nve.run(steps - 1)
    traj, epot, ekin = nve.frames()
    times = dt * np.arange(len(traj))
    if trajectory_output is not None:
        write_xyz_trajectory(trajectory_output, [init_mol.atom_symbol(i) for i in range(init_mol.natm)], traj[:, 0],
                             times)
    if energy_output is not None:
        write_md_energies(energy_output, times, epot[:, 0], ekin[:, 0])
    return traj[:, 0]


# ---- the active-learning loop (evcont/MD_utils.py:128-502) ------------------------------------------

def predict_energies(init_mol, geometries, one_rdm, two_rdm, overlap):
    """Continuation energies (with nuclear repulsion) at ``geometries`` (G, natm, 3), bohr: the
    reference's per-frame loop ``[approximate_ground_state_OAO(init_mol.copy().set_geom_(g), ...)[0]
    for g in trajectory]`` (evcont/MD_utils.py:271-282, :448-458) as ONE batched device step."""
    from .stackcache import as_device_stack
    stack = as_device_stack(one_rdm, two_rdm, overlap)
    eng = stack.engine
    geometries = np.ascontiguousarray(geometries, dtype=np.float64).reshape(-1, init_mol.natm, 3)
    ao = eng.ao_integrals(init_mol.sbasis(eng), geometries)
    return eng.energies(stack, ao)[0].cpu().numpy()   # energies only: no predicted RDMs, no gradient kernels


def oao_hamiltonian_rows(init_mol, geometries):
    """Device tensor (G, n^2 + n^4): ``h1`` then ``h2`` in the OAO basis for every geometry
    (``get_integrals(mol, get_basis(mol))``, evcont/MD_utils.py:383-392, batched: K9, K3, K4)."""
    import torch
    from .engine import get_engine
    eng = get_engine()
    geometries = np.ascontiguousarray(geometries, dtype=np.float64).reshape(-1, init_mol.natm, 3)
    ao = eng.ao_integrals(init_mol.sbasis(eng), geometries)
    x, _, _ = eng.loewdin(ao.ovlp)
    h1, h2, _ = eng.ao2oao(ao.hcore, ao.eri, x)
    G = geometries.shape[0]
    return torch.cat([h1.reshape(G, -1), h2.reshape(G, -1)], dim=1).contiguous()


def farthest_point_ham(init_mol, trn_geometries, trajectory):
    """Index of the trajectory frame whose OAO Hamiltonian is farthest (in the reference's metric
    ``|dh1|^2 + |dh2|^2 / 2``, evcont/MD_utils.py:394-405) from its nearest training Hamiltonian;
    first maximum, as the reference's strict ``>`` scan keeps it.  Returns ``(index, distances)``."""
    from ._lib import check
    from .engine import _ptr, get_engine
    eng = get_engine()
    n = int(init_mol.nao)
    rows_t = oao_hamiltonian_rows(init_mol, trn_geometries)
    rows_f = oao_hamiltonian_rows(init_mol, trajectory)
    dmin = eng.empty(rows_f.shape[0])
    eng._bind_stream()
    check(eng.lib.evc_min_sqdist(eng._ctx, rows_f.shape[0], rows_t.shape[0], n * n, rows_f.shape[1],
                                 _ptr(rows_f), _ptr(rows_t), _ptr(dmin)))
    d = dmin.cpu().numpy()
    return int(np.argmax(d)), d


def _save_stack(obj, i, trn_times, prune, workdir):
    sfx = "_{}".format(i) if prune else ""
    np.save(os.path.join(workdir, "overlap{}.npy".format(sfx)), obj.overlap)
    np.save(os.path.join(workdir, "one_rdm{}.npy".format(sfx)), obj.one_rdm)
    np.save(os.path.join(workdir, "two_rdm{}.npy".format(sfx)), obj.two_rdm)
    if trn_times is not None:
        np.savetxt(os.path.join(workdir, "trn_times{}.txt".format(sfx)), np.array(trn_times))


def _run_trajectory(obj, init_mol, i, steps, dt, workdir):
    traj = get_trajectory(init_mol.copy(), obj.overlap, obj.one_rdm, obj.two_rdm, steps=steps, dt=dt,
                          trajectory_output=os.path.join(workdir, "traj_EVCont_{}.xyz".format(i)),
                          energy_output=os.path.join(workdir, "ens_EVCont_{}.xyz".format(i)))
    np.save(os.path.join(workdir, "traj_EVCont_{}.npy".format(i)), traj)
    ens = np.ascontiguousarray(read_md_energies(os.path.join(workdir, "ens_EVCont_{}.xyz".format(i)))[:, 1])
    return traj, ens


def _prune(obj, init_mol, trajectory, updated_ens, trn_times, convergence_thresh):
    """Greedy removal of training points that do not change the trajectory energies by more than
    the threshold (evcont/MD_utils.py:286-313)."""
    keep = np.ones(len(trn_times), dtype=bool)
    for j in range(len(trn_times)):
        test_keep = keep.copy()
        test_keep[j] = False
        if np.sum(test_keep) >= 1:
            ids = np.ix_(test_keep, test_keep)
            ens = predict_energies(init_mol, trajectory, obj.one_rdm[ids], obj.two_rdm[ids], obj.overlap[ids])
            if np.all(abs(ens - updated_ens) < convergence_thresh):
                keep = test_keep
    keep_ids = np.nonzero(keep)[0]
    obj.prune_datapoints(keep_ids)
    return [trn_times[j] for j in keep_ids]


def converge_EVCont_MD(EVCont_obj, init_mol, steps=100, dt=1, convergence_thresh=1.0e-3,
                       prune_irrelevant_data=False, trn_times=[], data_addition="farthest_point_ham",
                       workdir=".", max_iterations=None):
    """On-the-fly training of the continuation along its own MD trajectories
    (evcont/MD_utils.py:128-502), same arguments, files (``overlap / one_rdm / two_rdm[_i].npy``,
    ``traj_EVCont_i.xyz / .npy``, ``ens_EVCont_i.xyz``, ``en_diff_i.txt``, ``trn_times[_i].txt``) and
    stopping rule: iterate {trajectory with the current training set; energies of its frames with the
    previous training set; stop once the largest difference stayed below ``convergence_thresh`` twice
    in a row; otherwise add the frame chosen by ``data_addition`` ("farthest_point_ham", "farthest_point"
    or "energy")}.  A non-empty ``trn_times`` resumes a previous run from the files in ``workdir``.

    ``init_mol`` is an :class:`evcont_b200.mol.MolLite` (``MolLite.from_mol`` makes one from a PySCF-like Mole with a
    named basis); one process (the reference's MPI rank 0 does
    all of this work and broadcasts).  ``workdir`` and ``max_iterations`` (a cap on added training
    points, ``None`` = the reference's unbounded loop) are additions.  Returns the last trajectory.
    """
    trn_times = list(trn_times)
    path = lambda name: os.path.join(workdir, name)
    if len(trn_times) < 1:
        i = 0
        trn_times = [0]
        EVCont_obj.append_to_rdms(init_mol.copy())
        _save_stack(EVCont_obj, i, None, prune_irrelevant_data, workdir)
        trajectory, updated_ens = _run_trajectory(EVCont_obj, init_mol, i, steps, dt, workdir)
        reference_ens = updated_ens[0]
        converged = False
    else:
        i = len(trn_times) - 1
        _save_stack(EVCont_obj, i, trn_times, prune_irrelevant_data, workdir)
        if os.path.exists(path("traj_EVCont_{}.npy".format(i))):
            trajectory = np.load(path("traj_EVCont_{}.npy".format(i)))
            updated_ens = np.ascontiguousarray(
                read_md_energies(path("ens_EVCont_{}.xyz".format(i)))[:, 1])
        else:
            trajectory, updated_ens = _run_trajectory(EVCont_obj, init_mol, i, steps, dt, workdir)
        if i > 0:
            reference_ens = predict_energies(init_mol, trajectory, EVCont_obj.one_rdm[:-1, :-1],
                                             EVCont_obj.two_rdm[:-1, :-1], EVCont_obj.overlap[:-1, :-1])
        else:
            reference_ens = updated_ens[0]
        if prune_irrelevant_data:
            trn_times = _prune(EVCont_obj, init_mol, trajectory, updated_ens, trn_times, convergence_thresh)
        converged = False
        if i >= 1:
            en_diff = np.loadtxt(path("en_diff_{}.txt".format(i - 1)))
            if np.max(en_diff) <= convergence_thresh:
                converged = True

    added = 0
    while True:
        en_diff = np.atleast_1d(abs(reference_ens - updated_ens))
        np.savetxt(path("en_diff_{}.txt".format(i)), np.array(en_diff))
        i += 1
        if converged and max(en_diff) <= convergence_thresh:
            break
        converged = bool(max(en_diff) <= convergence_thresh)
        if max_iterations is not None and added >= max_iterations:
            break

        if data_addition == "energy":
            trn_time = int(np.argmax(en_diff))
        elif data_addition in ("farthest_point", "farthest_point_ham"):
            trajs = [np.load(path("traj_EVCont_{}.npy".format(k))) for k in range(len(trn_times))]
            trn_geometries = [trajs[0][0]] + [trajs[k][trn_times[k + 1]] for k in range(len(trajs) - 1)]
            if data_addition == "farthest_point":
                trn_time = int(np.argmax(np.min(np.array(
                    [np.sum(abs(g - trajectory) ** 2, axis=(-1, -2)) for g in trn_geometries]), axis=0)))
            else:
                trn_time, _ = farthest_point_ham(init_mol, np.array(trn_geometries), trajectory)
        else:
            assert False
        trn_times.append(trn_time)
        EVCont_obj.append_to_rdms(init_mol.copy().set_geom_(trajectory[trn_time]))
        added += 1
        _save_stack(EVCont_obj, i, trn_times, prune_irrelevant_data, workdir)
        trajectory, updated_ens = _run_trajectory(EVCont_obj, init_mol, i, steps, dt, workdir)
        reference_ens = predict_energies(init_mol, trajectory, EVCont_obj.one_rdm[:-1, :-1],
                                         EVCont_obj.two_rdm[:-1, :-1], EVCont_obj.overlap[:-1, :-1])
        if prune_irrelevant_data:
            trn_times = _prune(EVCont_obj, init_mol, trajectory, updated_ens, trn_times, convergence_thresh)
    return trajectory
