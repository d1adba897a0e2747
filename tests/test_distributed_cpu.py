"""World-size-2 gloo tests of the multi-GPU host logic (pair sharding + all_gather
of the t-RDM stack slabs, geometry sharding) with the CPU oracle standing in for
the GPU kernel as ``pair_fn``.  CPU-only."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, random_civec


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ntrain, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from evcont_b200 import distributed as evd
    from oracle import trans_rdm as otr
    norb, nelec = 4, (2, 2)
    vecs = np.stack([random_civec(6, 6, 50 + k) for k in range(ntrain)])
    calls = []

    def pair_fn(v, plist):
        calls.append(list(plist))
        res = [otr.trans_rdm12(v[a], v[b], norb, nelec) for a, b in plist]
        ov = torch.tensor([(v[a] * v[b]).sum() for a, b in plist])
        return ov, torch.from_numpy(np.array([r[0] for r in res])), torch.from_numpy(np.array([r[1] for r in res]))

    S, one, two = evd.build_stack_sharded(vecs, norb, nelec, pair_fn=pair_fn)
    npairs = ntrain * (ntrain + 1) // 2
    lo, hi = evd.shard_range(npairs, rank, world)
    # the rank's share first, padded to the common slab size with repeats of the first pair
    slab = evd.slab_size(npairs, world)
    assert calls[0] == evd.tril_pairs(ntrain)[lo:hi] + [evd.tril_pairs(ntrain)[0]] * (slab - (hi - lo))
    idx = evd.shard_geometries(7, rank, world)
    E = torch.from_numpy(idx.astype(np.float64))
    G = E.reshape(-1, 1, 1).expand(-1, 3, 3).contiguous()
    Ea, Ga = evd.gather_predictions(E, G, 7)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), S=S.numpy(), one=one.numpy(), two=two.numpy(),
             Ea=Ea.numpy(), Ga=Ga.numpy())
    dist.destroy_process_group()


def _run(ntrain, tmp_path, world=2):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ntrain, str(tmp_path)), nprocs=world, join=True)
    return [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]


def test_sharded_stack_build_matches_serial(tmp_path):
    from oracle import trans_rdm as otr
    ntrain = 3
    r0, r1 = _run(ntrain, tmp_path)
    for k in ("S", "one", "two", "Ea", "Ga"):
        assert np.array_equal(r0[k], r1[k])
    vecs = [random_civec(6, 6, 50 + k) for k in range(ntrain)]
    for a in range(ntrain):
        for b in range(ntrain):
            hi, lo = max(a, b), min(a, b)
            d1, d2 = otr.trans_rdm12(vecs[hi], vecs[lo], 4, (2, 2))
            assert np.array_equal(r0["one"][a, b], d1)  # mirror blocks untransposed
            assert np.array_equal(r0["two"][a, b], d2)
            assert r0["S"][a, b] == (vecs[hi] * vecs[lo]).sum()
    assert np.array_equal(r0["Ea"], np.arange(7.0))
    assert np.array_equal(r0["Ga"][:, 2, 1], np.arange(7.0))


def test_single_state_leaves_one_rank_idle(tmp_path):
    r0, r1 = _run(1, tmp_path)  # one pair, two ranks: rank 0 has no work
    assert np.array_equal(r0["two"], r1["two"]) and r0["S"].shape == (1, 1)


def test_shard_ranges_cover_everything():
    from evcont_b200 import distributed as evd
    for n in (0, 1, 5, 210, 211):
        for w in (1, 2, 3, 8):
            spans = [evd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
            assert evd.slab_size(n, w) == max(h - l for l, h in spans)


# ---- e3: the pair-sharded prediction step (host logic: slab ranges, all_gather order, pair weights) ----
def _sharded_step_worker(rank, world, port, ntrain, norb, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import synthetic_stack
    from evcont_b200 import distributed as evd
    from oracle import subspace as osub
    ovlp, one, two = synthetic_stack(norb, ntrain, 5, 2)           # (N(N+1)/2, n2(n2+1)/2) rows
    rng = np.random.default_rng(1)
    h1 = rng.standard_normal((norb, norb)); h1 = h1 + h1.T
    h2 = rng.standard_normal((norb * norb, norb * norb)); h2 = (h2 + h2.T).reshape((norb,) * 4)
    npairs = two.shape[0]
    lo, hi = evd.shard_range(npairs, rank, world)
    rows = two[lo:hi]                                              # this rank's slab; numpy stands in for the row kernels
    hv = osub.compress_exchange(h2, 0.5)
    part = torch.from_numpy(rows @ hv).reshape(-1, 1)
    h_two = evd.gather_pair_entries(part, npairs).numpy()[:, 0]
    il = np.tril_indices(ntrain)
    H = np.zeros((ntrain, ntrain))
    H[il] = np.einsum("abij,ij->ab", one, h1)[il] + h_two
    H = np.tril(H) + np.tril(H, -1).T
    E, vec = osub._solve(H, ovlp, True)
    c = vec[:, 0] if vec.ndim == 2 else vec
    a, b = il[0][lo:hi], il[1][lo:hi]
    w = c[a] * c[b] * np.where(a == b, 1.0, 2.0)
    g2c = torch.from_numpy(w @ rows)
    dist.all_reduce(g2c)
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), H=H, g2c=g2c.numpy(), c=c, h1=h1, h2=h2)
    dist.destroy_process_group()


def test_pair_sharded_step_matches_the_unsharded_formulas(tmp_path):
    from conftest import synthetic_stack
    from oracle import gradients as og
    from oracle import subspace as osub
    ntrain, norb, world = 5, 3, 2
    mp.spawn(_sharded_step_worker, args=(world, _free_port(), ntrain, norb, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(tmp_path, f"s{k}.npz")) for k in range(world)]
    for k in ("H", "g2c", "c"):
        assert np.array_equal(r[0][k], r[1][k]), k
    ovlp, one, two = synthetic_stack(norb, ntrain, 5, 2)
    Href = osub.subspace_hamiltonian(r[0]["h1"], r[0]["h2"], one, two)
    assert np.abs(np.tril(r[0]["H"]) - np.tril(Href)).max() < 1e-12
    _, Gamma = og.predict_rdms(r[0]["c"], one, two, norb)
    assert np.abs(osub.restore_exchange(r[0]["g2c"], norb) - Gamma).max() < 1e-12
