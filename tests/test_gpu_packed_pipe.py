"""The persistent warp-specialised K4p / K8a kernels (csrc/packed_pipe.cu), the packed AO two-electron
arrays (erip / eri_ip1p) and the packed streaming gradient, through the C ABI.

Correctness against the reference-generated goldens is covered by tests/test_gpu_predict.py (every
PREDICT_CASE has n <= 10 and therefore runs these kernels); here: batches large enough that every
CTA walks its three-stage ring several times, bit-identity between the packed and the full-tensor
inputs, the packed integral output of K9, and the 11..13-orbital kernels on the same layouts."""
import numpy as np
import pytest

from conftest import synthetic_stack

pytestmark = pytest.mark.gpu


def _setup(norb, natm, ntrain, layout, nmol, seed=900):
    from evcont_b200.engine import DeviceAO
    from evcont_b200.mol import ao_bundle, synthetic_mol
    from evcont_b200.stackcache import as_device_stack
    ovlp, one, two = synthetic_stack(norb, ntrain, seed, layout)
    stack = as_device_stack(one, two, ovlp)
    mols = [synthetic_mol(norb, natm, seed=seed + 1 + k) for k in range(nmol)]
    bundles = [ao_bundle(m) for m in mols]
    return stack, mols, bundles, (ovlp, one, two), DeviceAO


@pytest.mark.parametrize("norb,natm,ntrain", [(10, 10, 6), (6, 6, 3), (7, 3, 4), (3, 2, 2)])
def test_ring_wraps_and_every_copy_is_bit_identical(norb, natm, ntrain):
    """7 distinct geometries tiled to 1500: each CTA of the persistent kernels handles ~10 geometries
    (more than three ring revolutions); equal geometries must give bit-identical results wherever they
    sit in the batch, and agree with a 7-geometry call (one geometry per CTA, no ring reuse)."""
    stack, mols, bundles, _, DeviceAO = _setup(norb, natm, ntrain, 6, 7)
    eng = stack.engine
    small = DeviceAO.from_bundles(eng, bundles)
    E0, g0, _, _, c0 = eng.energy_with_grad(stack, small)
    big = DeviceAO.from_bundles(eng, [bundles[k % 7] for k in range(1500)])
    E1, g1, _, _, c1 = eng.energy_with_grad(stack, big)
    E1, g1 = E1.cpu().numpy(), g1.cpu().numpy()
    for k in range(7, 1500):   # bit-identical wherever a geometry sits in the batch
        assert E1[k] == E1[k % 7], k
        assert np.array_equal(g1[k], g1[k % 7]), k
    # (the small batch takes the streaming forms of K5p / K7p: same values to rounding, not the same bits)
    assert np.abs(E1[:7] - E0.cpu().numpy()).max() < 1e-11
    assert np.abs(g1[:7] - g0.cpu().numpy()).max() < 1e-10


@pytest.mark.parametrize("norb,natm,ntrain", [(10, 10, 6), (5, 3, 3), (12, 4, 3), (13, 3, 4)])
def test_against_oracle_and_full_step(norb, natm, ntrain):
    """Energies / forces of the packed step vs the numpy oracle of the reference (1e-10 Ha, 1e-8 Ha/bohr),
    including the 11..13-orbital kernels that read the same packed layouts."""
    from oracle import gradients as og
    stack, mols, bundles, (ovlp, one, two), DeviceAO = _setup(norb, natm, ntrain, 6, 3, seed=321)
    eng = stack.engine
    ao = DeviceAO.from_bundles(eng, bundles)
    E, g, _, _, _ = eng.energy_with_grad(stack, ao)
    for k, m in enumerate(mols):
        e_ref, g_ref = og.get_energy_with_grad(m, one, two, ovlp)
        assert abs(E[k].item() - e_ref) < 1e-10
        assert np.abs(g[k].cpu().numpy() - g_ref).max() < 1e-8


@pytest.mark.parametrize("norb,natm,ntrain", [(10, 10, 6), (8, 4, 3), (13, 3, 4)])
def test_packed_inputs_equal_full_inputs(norb, natm, ntrain):
    """erip / eri_ip1p handed in directly (evc_ao_pack8 on the device, pack_ao_host on the host) give the
    same bits as the full tensors (which the step packs itself)."""
    from evcont_b200.engine import HostAO, pack_ao_host
    stack, mols, bundles, _, DeviceAO = _setup(norb, natm, ntrain, 5, 9, seed=77)
    eng = stack.engine
    ao = DeviceAO.from_bundles(eng, bundles)
    E0, g0, _, _, _ = eng.energy_with_grad(stack, ao)
    aop = ao.to_packed()
    assert aop.packed and aop.eri is None and aop.erip.shape[1] == norb * (norb + 1) // 2
    E1, g1, _, _, _ = eng.energy_with_grad(stack, aop)
    assert np.array_equal(E0.cpu().numpy(), E1.cpu().numpy())
    assert np.array_equal(g0.cpu().numpy(), g1.cpu().numpy())
    # host packing agrees with the device packing
    erip_h, ip1p_h = pack_ao_host(np.stack([b["eri"] for b in bundles]), np.stack([b["eri_ip1"] for b in bundles]))
    npair = norb * (norb + 1) // 2
    assert np.array_equal(erip_h[:, :, :npair], aop.erip.cpu().numpy()[:, :, :npair])
    assert np.array_equal(ip1p_h, aop.eri_ip1p.cpu().numpy())
    # the host-buffer entry point with packed host arrays
    hao = HostAO.from_bundles(bundles, packed=True)
    Eh, gh = eng.energy_with_grad_host(stack, hao, chunk=4)
    assert np.array_equal(Eh.numpy(), E0.cpu().numpy())
    assert np.array_equal(gh.numpy(), g0.cpu().numpy())
    with pytest.raises(ValueError):
        eng.energy_with_grad(stack, aop, want_rdms=True)


def test_packed_integral_output_of_k9():
    """evc_ao_integrals_s_packed writes exactly the entries evc_ao_pack8 picks from the full arrays."""
    from evcont_b200.engine import get_engine
    eng = get_engine()
    rng = np.random.default_rng(5)
    co = np.zeros((40, 10, 3))
    co[:, :, 0] = 1.78596 * np.arange(10)
    co += 0.2 * rng.standard_normal(co.shape)
    sb = eng.sbasis(["H"] * 10, "sto-6g")
    full = eng.ao_integrals(sb, co)
    pk = eng.ao_integrals(sb, co, packed=True)
    ref = full.to_packed()
    npair = 55
    assert pk.packed and pk.erip.shape == (40, npair, eng.lib.evc_erip_pitch(10))
    assert np.array_equal(pk.erip.cpu().numpy()[:, :, :npair], ref.erip.cpu().numpy()[:, :, :npair])
    assert np.array_equal(pk.eri_ip1p.cpu().numpy(), ref.eri_ip1p.cpu().numpy())
    for k in ("ovlp", "hcore", "ipovlp", "hcore_deriv", "e_nuc", "grad_nuc"):
        assert np.array_equal(getattr(pk, k).cpu().numpy(), getattr(full, k).cpu().numpy()), k
    # one geometry (quartets split over several CTAs) gives the same bits
    one = eng.ao_integrals(sb, co[:1], packed=True)
    assert np.array_equal(one.erip.cpu().numpy()[0, :, :npair], pk.erip.cpu().numpy()[0, :, :npair])
    assert np.array_equal(one.eri_ip1p.cpu().numpy()[0], pk.eri_ip1p.cpu().numpy()[0])


def test_step_from_coordinates_uses_packed_arrays():
    from evcont_b200.engine import get_engine
    from evcont_b200.stackcache import as_device_stack
    eng = get_engine()
    ovlp, one, two = synthetic_stack(10, 5, 12, 6)
    stack = as_device_stack(one, two, ovlp)
    rng = np.random.default_rng(6)
    co = np.zeros((300, 10, 3))
    co[:, :, 0] = 1.78596 * np.arange(10)
    co += 0.2 * rng.standard_normal(co.shape)
    sb = eng.sbasis(["H"] * 10, "sto-6g")
    E1, g1, _, _, _ = eng.energy_with_grad_coords(stack, sb, co)                 # packed integrals
    E0, g0, _, _, _ = eng.energy_with_grad(stack, eng.ao_integrals(sb, co))      # full tensors, packed by the step
    assert np.array_equal(E0.cpu().numpy(), E1.cpu().numpy())
    assert np.array_equal(g0.cpu().numpy(), g1.cpu().numpy())


@pytest.mark.parametrize("norb", list(range(2, 11)))
def test_k4p_every_orbital_count_repeated(norb):
    """K4p alone (hvec = [h1 | tril h2p] and the T image), pipelined kernel vs the one-CTA-per-geometry
    kernel of packed.cu, for EVERY template instance n = 2..10, one geometry per CTA (3) and several ring
    revolutions (700), each launched repeatedly: a scheduling-dependent fault of one instance (n = 5 once
    dropped a k-step of one tile in most launches) does not hide behind a single lucky run."""
    import ctypes as C
    import torch
    from evcont_b200.engine import get_engine, _ptr
    from evcont_b200.mol import ao_bundle, synthetic_mol
    eng = get_engine()
    lib = eng.lib
    fn = lib.evc_debug_packed_ao2oao
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int]
    fn.restype = C.c_int
    b = [ao_bundle(synthetic_mol(norb, 2, seed=10 + k)) for k in range(8)]
    L8, el = int(lib.evc_packed_row_len(norb)), int(lib.evc_erip_len(norb))
    for G, reps in ((3, 12), (700, 4)):
        S = eng.to_device(np.stack([b[k % 8]["ovlp"] for k in range(G)]))
        X, _, _ = eng.loewdin(S)
        hc = eng.to_device(np.stack([b[k % 8]["hcore"] for k in range(G)]))
        erip, _ = eng.ao_pack8(eng.to_device(np.stack([b[k % 8]["eri"] for k in range(G)])), None)

        def run(pipe):
            hv = torch.zeros(G, L8, dtype=torch.float64, device=eng.device)
            T = torch.zeros(G, el, dtype=torch.float64, device=eng.device)
            eng._bind_stream()
            assert fn(eng._ctx, G, norb, _ptr(X), _ptr(hc), _ptr(erip), _ptr(hv), _ptr(T), pipe) == 0
            torch.cuda.synchronize()
            npk, pitch = norb * (norb + 1) // 2, int(lib.evc_erip_pitch(norb))
            return hv.cpu().numpy(), T.cpu().numpy().reshape(G, npk, pitch)[:, :, :npk]

        hv0, T0 = run(0)
        scale = max(1.0, np.abs(hv0).max())
        for _ in range(reps):
            hv1, T1 = run(1)
            assert np.abs(hv1 - hv0).max() < 1e-11 * scale
            assert np.abs(T1 - T0).max() < 1e-11 * scale


@pytest.mark.parametrize("norb", list(range(2, 11)))
def test_step_every_orbital_count_against_oracle(norb):
    """The whole packed step (every template instance of K4p AND K8a, n = 2..10) against the numpy oracle of the
    reference, one geometry per CTA and with the rings wrapping, three launches each."""
    from oracle import gradients as og
    natm = 2 if norb < 4 else 3
    stack, mols, bundles, (ovlp, one, two), DeviceAO = _setup(norb, natm, 3, 6, 3, seed=4321 + norb)
    eng = stack.engine
    ref = [og.get_energy_with_grad(m, one, two, ovlp) for m in mols]
    for G in (3, 600):
        ao = DeviceAO.from_bundles(eng, [bundles[k % 3] for k in range(G)]).to_packed()
        for _ in range(3):
            E, g, _, _, _ = eng.energy_with_grad(stack, ao)
            E, g = E.cpu().numpy(), g.cpu().numpy()
            for k in range(G):
                assert abs(E[k] - ref[k % 3][0]) < 1e-10, (G, k)
                assert np.abs(g[k] - ref[k % 3][1]).max() < 1e-8, (G, k)
