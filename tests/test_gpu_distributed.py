"""The NCCL stack build on real GPUs: pairs sharded over 2 ranks, ONE all_gather, placement kernel -- must equal
the single-GPU build bit for bit ("deterministic across GPU counts").  Needs >= 2 GPUs on the box
(skipped otherwise); the host logic alone is covered on CPU by tests/test_distributed_cpu.py."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, random_civec

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ntrain, norb, nocc, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from math import comb
    from evcont_b200 import distributed as evd
    from evcont_b200.engine import get_engine
    na = comb(norb, nocc)
    vecs = np.stack([random_civec(na, na, 70 + k, symmetric=True) for k in range(ntrain)])
    eng = get_engine(torch.device("cuda", rank))
    vd = eng.to_device(vecs)
    S, one, two = evd.build_stack_sharded(vd, norb, (nocc, nocc), device=torch.device("cuda", rank))
    S1, one1, two1 = evd.build_stack_single(vd, norb, (nocc, nocc), device=torch.device("cuda", rank))
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), S=S.cpu().numpy(), one=one.cpu().numpy(), two=two.cpu().numpy(),
             S1=S1.cpu().numpy(), one1=one1.cpu().numpy(), two1=two1.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("ntrain,norb,nocc", [(5, 6, 3), (7, 8, 4), (6, 10, 5)])
def test_nccl_sharded_build_equals_single_gpu_bitwise(tmp_path, ntrain, norb, nocc):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), ntrain, norb, nocc, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(tmp_path, f"r{k}.npz")) for k in range(world)]
    for k in ("S", "one", "two"):
        assert np.array_equal(r[0][k], r[1][k]), k                 # identical on every rank
        assert np.array_equal(r[0][k], r[0][k + "1"]), k           # and bit-identical to the single-GPU build
    # mirror blocks are the untransposed copies (evcont/FCI_EVCont.py:124-127)
    assert np.array_equal(r[0]["two"][0, ntrain - 1], r[0]["two"][ntrain - 1, 0])
    assert np.array_equal(r[0]["S"], r[0]["S"].T)


def test_single_gpu_rows_and_scatter_against_batch_entry():
    """evc_trans_rdm12_batch_strided + evc_stack_scatter_rows give the stack FCI_EVCont_obj.from_civecs builds
    from the plain batch entry."""
    from evcont_b200 import distributed as evd
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    from evcont_b200.engine import get_engine
    norb, nocc, ntrain = 6, 3, 4
    vecs = [random_civec(20, 20, 90 + k, symmetric=True) for k in range(ntrain)]
    eng = get_engine()
    S, one, two = evd.build_stack_single(eng.to_device(np.stack(vecs)), norb, (nocc, nocc))
    ref = FCI_EVCont_obj.from_civecs(vecs, norb, (nocc, nocc))
    assert np.array_equal(S.cpu().numpy(), ref.overlap)
    assert np.array_equal(one.cpu().numpy(), ref.one_rdm)
    assert np.array_equal(two.cpu().numpy(), ref.two_rdm)


def test_a_share_of_the_pairs_has_the_bits_of_the_whole_build():
    """One GPU: the rows a rank computes for its share (planned for the pair count of the whole build,
    evc_trans_rdm12_plan_pairs) are bit-identical to the same rows of the full call -- at H10 sizes, where the
    alpha-slice count would otherwise follow the number of pairs in the call."""
    from math import comb
    from evcont_b200 import distributed as evd
    from evcont_b200.engine import get_engine
    eng = get_engine()
    ntrain, norb, nocc = 6, 10, 5
    na = comb(norb, nocc)
    vecs = eng.to_device(np.stack([random_civec(na, na, 70 + k, symmetric=True) for k in range(ntrain)]))
    pairs = evd.tril_pairs(ntrain)
    used = norb ** 4 + norb ** 2 + 1          # [dm2 | dm1 | ovlp]; the row is padded to an even length
    full = eng.trans_rdm12_rows(vecs, pairs, norb, (nocc, nocc)).cpu().numpy()[:, :used]
    for world in (2, 4, 8):
        for rank in range(world):
            lo, hi = evd.shard_range(len(pairs), rank, world)
            part = eng.trans_rdm12_rows(vecs, pairs[lo:hi], norb, (nocc, nocc), plan_pairs=len(pairs)).cpu().numpy()
            assert np.array_equal(part[:, :used], full[lo:hi]), (world, rank)
            if hi - lo < len(pairs) // 2:         # without the announcement the slice count follows the call
                own = eng.trans_rdm12_rows(vecs, pairs[lo:hi], norb, (nocc, nocc)).cpu().numpy()[:, :used]
                assert np.abs(own - full[lo:hi]).max() < 1e-12
