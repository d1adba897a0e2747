"""e3: one prediction step on a stack sharded by training pairs (evcont_b200/distributed.py,
PairShardedStack / sharded_energy_with_grad) against the unsharded device step:
on one GPU with the slabs held side by side (the arithmetic without the collectives), with per-pair directories
as the reference's Zundel run stores them, and over NCCL on two GPUs (skipped on a single-GPU box)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, synthetic_stack

pytestmark = pytest.mark.gpu


def _problem(norb, natm, ntrain, G, seed=31):
    from evcont_b200.engine import DeviceAO, get_engine
    from evcont_b200.mol import ao_bundle, synthetic_mol
    eng = get_engine()
    ovlp, one, two = synthetic_stack(norb, ntrain, seed, 2)
    ao = DeviceAO.from_bundles(eng, [ao_bundle(synthetic_mol(norb, natm, seed=seed + 1 + k)) for k in range(G)])
    return eng, ovlp, one, two, ao


@pytest.mark.parametrize("norb,natm,ntrain,G,nslab", [(6, 4, 5, 1, 3), (5, 3, 4, 3, 2), (7, 3, 6, 40, 4), (4, 2, 3, 2, 7)])
def test_local_slabs_equal_the_unsharded_step(norb, natm, ntrain, G, nslab):
    """Odd and even row lengths, the streaming (G <= 16) and the tensor-core forms of the row kernels, more slabs
    than some ranks have pairs (empty slabs)."""
    from evcont_b200 import distributed as evd
    from evcont_b200.stackcache import as_device_stack
    eng, ovlp, one, two, ao = _problem(norb, natm, ntrain, G)
    stack = as_device_stack(one, two, ovlp)
    E0, g0, gam0, Gam0, _ = eng.energy_with_grad(stack, ao, want_rdms=True)
    slabs = [evd.PairShardedStack.from_full(ovlp, one, two, r, nslab, engine=eng) for r in range(nslab)]
    E1, g1, gam1, Gam1 = evd.sharded_energy_with_grad(slabs, ao, want_rdms=True)
    assert np.abs((E1 - E0).cpu().numpy()).max() < 1e-10
    assert np.abs((g1 - g0).cpu().numpy()).max() < 1e-9
    assert np.abs((gam1 - gam0).cpu().numpy()).max() < 1e-10
    assert np.abs((Gam1 - Gam0).cpu().numpy()).max() < 1e-10


def test_slab_from_pair_directories(tmp_path):
    """The reference's on-disk form: one directory per training pair (04_Zundel_continuation_MD.py:99-128)."""
    from evcont_b200 import distributed as evd
    eng, ovlp, one, two, ao = _problem(6, 3, 4, 2)
    il = np.tril_indices(4)
    for p, (a, b) in enumerate(zip(*il)):
        d = tmp_path / f"MPS_cross_{a}_{b}"
        os.makedirs(d)
        np.save(d / "ovlp.npy", ovlp[a, b]); np.save(d / "one_rdm.npy", one[a, b]); np.save(d / "two_rdm.npy", two[p])
    ref = evd.sharded_energy_with_grad([evd.PairShardedStack.from_full(ovlp, one, two, 0, 1, engine=eng)], ao)
    slabs = [evd.PairShardedStack.from_pair_dirs(str(tmp_path), 4, r, 3, engine=eng) for r in range(3)]
    assert [s.hi - s.lo for s in slabs] == [3, 3, 4]
    out = evd.sharded_energy_with_grad(slabs, ao)
    assert np.abs((out[0] - ref[0]).cpu().numpy()).max() < 1e-11
    assert np.abs((out[1] - ref[1]).cpu().numpy()).max() < 1e-10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from evcont_b200 import distributed as evd
    from evcont_b200.engine import DeviceAO, get_engine
    from evcont_b200.mol import ao_bundle, synthetic_mol
    from evcont_b200.stackcache import as_device_stack
    eng = get_engine(torch.device("cuda", rank))
    ovlp, one, two = synthetic_stack(6, 5, 31, 2)
    ao = DeviceAO.from_bundles(eng, [ao_bundle(synthetic_mol(6, 4, seed=32 + k)) for k in range(2)])
    shard = evd.PairShardedStack.from_full(ovlp, one, two, rank, world, engine=eng)
    E, g = evd.sharded_energy_with_grad(shard, ao)
    E0, g0, _, _, _ = eng.energy_with_grad(as_device_stack(one, two, ovlp), ao)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), E=E.cpu().numpy(), g=g.cpu().numpy(), E0=E0.cpu().numpy(),
             g0=g0.cpu().numpy())
    dist.destroy_process_group()


def test_nccl_pair_sharded_step(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = [np.load(os.path.join(tmp_path, f"r{k}.npz")) for k in range(2)]
    assert np.array_equal(r[0]["E"], r[1]["E"]) and np.array_equal(r[0]["g"], r[1]["g"])   # identical on every rank
    assert np.abs(r[0]["E"] - r[0]["E0"]).max() < 1e-10 and np.abs(r[0]["g"] - r[0]["g0"]).max() < 1e-9
