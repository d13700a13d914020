"""Parity of the CUDA path (through the C ABI, libpcdgpu.so) against the golden vectors and the C++
oracle.  Everything here is bit-exact: field elements are canonical and points are compared after
normalisation to affine."""
import numpy as np
import pytest

import c_oracle as co
import codec
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import pcd_b200
    c = pcd_b200.Context(0)
    yield c
    c.close()


# ---- NTT ---------------------------------------------------------------------------------------
FLAVOURS = (("fft", 0, 0), ("coset_fft", 0, 1), ("ifft", 1, 0), ("coset_ifft", 1, 1))


def test_ntt_golden(ctx):
    for case in codec.load("ntt"):
        x = codec.hex_to_u64(case["input"], 5)
        for name, inv, cos in FLAVOURS:
            got = ctx.ntt(case["field"], x, inv, cos)
            assert codec.u64_to_hex(got) == case[name], (case["field"], case["log_n"], name)


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("log_n", [1, 2, 4, 8, 9, 10, 11, 12, 13, 15, 16, 17])
def test_ntt_vs_oracle(ctx, field, log_n):
    x = codec.random_field_elems(1 << log_n, field, 1000 + log_n)
    for name, inv, cos in FLAVOURS:
        got = ctx.ntt(field, x, inv, cos)
        ref = co.ntt(field, x, inv, cos, threads=8)
        assert np.array_equal(got, ref), (field, log_n, name)


@pytest.mark.parametrize("log_n", [18, 20, 22])
def test_ntt_large_r4(ctx, log_n):
    x = codec.random_field_elems(1 << log_n, 0, 77 + log_n)
    got = ctx.ntt(0, x, 0, 1)
    if log_n <= 20:
        assert np.array_equal(got, co.ntt(0, x, 0, 1, threads=8))
    # size-independent property: coset_ifft(coset_fft(x)) == x
    assert np.array_equal(ctx.ntt(0, got, 1, 1), x)


def test_ntt_domain_too_large(ctx):
    import pcd_b200
    x = np.zeros((4, 5), dtype=np.uint64)
    with pytest.raises(pcd_b200.PcdGpuError) as e:
        ctx._check(ctx.lib.pcdgpu_ntt(ctx.h, 1, x.ctypes.data, 18, 0, 0))
    assert e.value.code == -4


# ---- MSM ---------------------------------------------------------------------------------------
def test_msm_golden(ctx):
    for case in codec.load("msm"):
        cv = case["curve"]
        bases = codec.hex_to_u64(case["bases"], codec.POINT_LIMBS[cv])
        sc = codec.hex_to_u64(case["scalars"], 5)
        got = ctx.msm(cv, bases, sc)
        assert codec.u64_to_hex(got) == case["result"], (cv, case["n"])


def test_msm_empty(ctx):
    for cv in (0, 1, 2, 3):
        got = ctx.msm(cv, np.zeros((0, codec.POINT_LIMBS[cv]), np.uint64), np.zeros((0, 5), np.uint64))
        assert not got.any()


@pytest.mark.parametrize("curve,n", [(0, 100), (0, 1000), (0, 1 << 14), (2, 1000), (2, 1 << 13), (1, 300), (1, 4096),
                                     (3, 300), (3, 2048)])
@pytest.mark.parametrize("dist", ["U", "W"])
def test_msm_vs_oracle(ctx, curve, n, dist):
    pts = synth.random_points(n, curve, 5 + n)
    pts[3] = 0                      # a point at infinity among the bases
    pts[7] = pts[6]                 # duplicate base
    sc = synth.random_scalars(n, curve, 11 + n, dist)
    sc[7] = sc[6]                   # ... with the same scalar: equal points meet in every bucket
    sc[8] = codec.int_to_limbs(codec.CURVE_ORDER[curve] - 1)
    ref = co.msm(curve, pts, sc, threads=8)
    assert np.array_equal(ctx.msm(curve, pts, sc), ref)
    # every window size the kernel can be asked for gives the same affine point
    for c in (5, 9):
        ctx.set_msm_window(c)
        try:
            assert np.array_equal(ctx.msm(curve, pts, sc), ref), c
        finally:
            ctx.set_msm_window(0)


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_msm_resident_bases(ctx, curve):
    import pcd_b200
    n = 600 if curve in (1, 3) else 3000
    pts = synth.random_points(n, curve, 21)
    sc = synth.random_scalars(n, curve, 22, "W")
    ref = co.msm(curve, pts, sc, threads=8)
    ref_off = co.msm(curve, pts[5:], sc[: n - 5], threads=8)
    for pre in (False, True):
        b = pcd_b200.Bases(ctx, curve, pts, precompute=pre)
        assert np.array_equal(b.msm(sc), ref), pre
        assert np.array_equal(b.msm(sc[: n - 5], offset=5), ref_off), pre
        b.close()


def test_msm_heavy_bucket(ctx):
    # every scalar equal: one bucket per window receives all points (the heavy-bucket kernel)
    n = 5000
    pts = synth.random_points(n, 0, 33)
    sc = np.tile(codec.int_to_limbs(0x1F2E3D4C5B6A7988 * 2 ** 64 + 5), (n, 1))
    assert np.array_equal(ctx.msm(0, pts, sc), co.msm(0, pts, sc, threads=8))
    sc1 = np.tile(codec.int_to_limbs(1), (n, 1))
    assert np.array_equal(ctx.msm(0, pts, sc1), co.point_sum(0, pts))


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_fixed_base_mul(ctx, curve):
    n = 64
    k = synth.random_scalars(n, curve, 44)
    k[0] = 0
    k[1] = codec.int_to_limbs(1)
    g = synth.generator_limbs(curve)
    assert np.array_equal(ctx.fixed_base_mul(curve, g, k), co.fixed_base_mul(curve, g, k))


def test_msm_large_linearity(ctx):
    # BASELINE size 2^20 through a size-independent property: MSM(P, s) + MSM(P, t) == MSM(P, s + t)
    import torch
    n = 1 << 20
    dev = torch.device("cuda:0")
    k = torch.from_numpy(codec.random_field_elems(n, 0, 55).view(np.int64)).to(dev)
    pts = torch.empty((n, 10), dtype=torch.int64, device=dev)
    ctx.fixed_base_mul_dev(0, synth.generator_limbs(0), k.data_ptr(), n, pts.data_ptr())
    s = codec.random_field_elems(n, 0, 56)
    t = codec.random_field_elems(n, 0, 57)
    s[:, 4] >>= np.uint64(1)
    t[:, 4] >>= np.uint64(1)   # keep s + t below 2^298 and below r: no modular wrap needed
    si = s.astype(object)
    st = np.zeros_like(s)
    carry = np.zeros(n, dtype=object)
    for j in range(5):
        v = si[:, j] + t[:, j].astype(object) + carry
        st[:, j] = (v % (1 << 64)).astype(np.uint64)
        carry = v >> 64
    host_pts = pts.cpu().numpy().view(np.uint64)
    a = ctx.msm(0, host_pts, s)
    b = ctx.msm(0, host_pts, t)
    ab = ctx.msm(0, host_pts, st)
    assert np.array_equal(co.point_sum(0, np.stack([a, b])), ab)
    # and against the discrete logs: sum_i s_i k_i * G
    p = codec.FIELD_P[0]
    kk = codec.limbs_to_int  # noqa
    k_host = k.cpu().numpy().view(np.uint64)
    acc = 0
    for i in range(0, n, 1):
        if i >= 4096:
            break
        acc += codec.limbs_to_int(k_host[i]) * codec.limbs_to_int(s[i])
    part = ctx.msm(0, host_pts[:4096], s[:4096])
    assert np.array_equal(part, co.fixed_base_mul(0, synth.generator_limbs(0), codec.ints_to_limbs([acc % p]), 1)[0])


# ---- witness map and Groth16 ------------------------------------------------------------------------
def _index(ctx, inst, precompute=False):
    import pcd_b200
    pk = pcd_b200.ProvingKey(pairing=inst["pairing"], **inst["pk"])
    cm = pcd_b200.ConstraintMatrices(inst["pairing"], inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"],
                                     inst["C"])
    g = pcd_b200.Groth16(ctx, inst["pairing"])
    return g, g.index(pk, cm, precompute)


def test_groth16_golden(ctx):
    for case in codec.load("groth16"):
        pid = case["pairing"]
        g1, g2 = codec.G1_OF[pid], codec.G2_OF[pid]
        pk = {k: codec.hex_to_u64(v, codec.POINT_LIMBS[g2 if "g2" in k else g1]) for k, v in case["pk"].items()}
        for k in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2"):
            pk[k] = pk[k][0]
        inst = dict(pairing=pid, pk=pk, num_inputs=case["num_inputs"], num_witness=case["num_witness"],
                    A=codec.csr_from_golden(case["A"]), B=codec.csr_from_golden(case["B"]),
                    C=codec.csr_from_golden(case["C"]))
        z = codec.hex_to_u64(case["z"], 5)
        for pre in (False, True):
            g, idx = _index(ctx, inst, pre)
            assert codec.u64_to_hex(g.witness_map(idx, z)) == case["h"]
            proof = g.create_proof_with_reduction(idx, z, codec.hex_to_u64(case["r"]), codec.hex_to_u64(case["s"]))
            assert codec.u64_to_hex(proof.affine_limbs()) == case["proof_affine"]
            assert g.serialize(proof).hex() == case["proof_bytes"]
            idx.close()


@pytest.mark.parametrize("pairing,m", [(0, 1000), (1, 1000), (0, (1 << 13) - 2), (1, (1 << 12) - 2)])
def test_groth16_vs_oracle_and_trapdoor(ctx, pairing, m):
    inst = synth.make_instance(pairing, m, seed=900 + m, bitlike=0.4)
    p = codec.FIELD_P[pairing]
    r, s = pow(3, 200, p), pow(7, 190, p)
    rl, sl = codec.int_to_limbs(r), codec.int_to_limbs(s)
    g, idx = _index(ctx, inst)
    h = g.witness_map(idx, inst["z"])
    assert np.array_equal(h, co.witness_map(pairing, inst["A"], inst["B"], inst["C"], m, inst["num_inputs"], inst["z"], 8))
    proof = g.create_proof_with_reduction(idx, inst["z"], rl, sl)
    ref = co.groth16_prove(pairing, inst["pk"], inst["A"], inst["B"], inst["C"], m, inst["num_inputs"],
                           inst["num_witness"], inst["z"], rl, sl, threads=8)
    assert np.array_equal(proof.affine_limbs(), ref)
    assert np.array_equal(proof.affine_limbs(), synth.trapdoor_proof(inst, r, s))
    assert g.serialize(proof) == co.serialize_proof(pairing, ref)
    idx.close()
    g2, idx2 = _index(ctx, inst, precompute=True)
    proof2 = g2.create_proof_with_reduction(idx2, inst["z"], rl, sl)
    assert np.array_equal(proof2.affine_limbs(), ref)
    # the reference's rng contract: prove() draws r first, then s
    draws = iter([rl, sl])
    proof3 = g2.prove(idx2, inst["z"], lambda f: next(draws))
    assert np.array_equal(proof3.affine_limbs(), ref)
    idx2.close()


def test_groth16_domain_limit(ctx):
    # helper-side field q4: radix 2 up to 2^17, mixed radix up to 49 * 2^17; anything larger must be refused,
    # not mis-computed
    import pcd_b200
    m = (49 << 17) + 1
    ptr = np.zeros(m + 1, dtype=np.uint32)
    empty = (ptr, np.zeros(0, np.uint32), np.zeros((0, 5), np.uint64))
    cm = pcd_b200.ConstraintMatrices(1, 1, 0, empty, empty, empty)
    import ctypes
    h = ctypes.c_void_p()
    args = []
    for (p_, c_, v_) in (cm.a, cm.b, cm.c):
        args += [p_.ctypes.data, c_.ctypes.data, v_.ctypes.data]
    ctx._check(ctx.lib.pcdgpu_r1cs_upload(ctx.h, 1, m, 1, 0, *args, ctypes.byref(h)))
    z = np.zeros((1, 5), dtype=np.uint64)
    out = np.zeros((ctx.lib.pcdgpu_r1cs_domain_size(h), 5), dtype=np.uint64)
    rc = ctx.lib.pcdgpu_witness_map(ctx.h, h, z.ctypes.data, out.ctypes.data)
    assert rc == -4
    ctx.lib.pcdgpu_r1cs_free(h)


# ---- mixed-radix domains (q4 = MNT6-298 Fr beyond 2^17) ------------------------------------------------
@pytest.mark.parametrize("a,b", [(1, 0), (2, 0), (1, 3), (2, 2), (1, 10), (2, 9), (1, 13), (2, 12)])
def test_ntt_mixed_radix_vs_oracle(ctx, a, b):
    n = (7 ** a) << b
    x = codec.random_field_elems(n, 1, 300 + 10 * a + b)
    for name, inv, cos in FLAVOURS:
        got = ctx.ntt_general(1, x, a, b, inv, cos)
        ref = co.ntt_general(1, x, a, b, inv, cos, threads=8)
        assert np.array_equal(got, ref), (a, b, name)
    # round trip on the coset
    assert np.array_equal(ctx.ntt_general(1, ctx.ntt_general(1, x, a, b, False, True), a, b, True, True), x)


def test_mixed_radix_refused_on_r4(ctx):
    import pcd_b200
    x = np.zeros((7, 5), dtype=np.uint64)
    with pytest.raises(pcd_b200.PcdGpuError) as e:
        ctx.ntt_general(0, x, 1, 0)
    assert e.value.code == -4


def test_groth16_mixed_radix_domain(ctx):
    """helper side (MNT6-298, Fr = q4): 2^17 + 1 constraints need the domain 49 * 2^12 = 200704"""
    import pcd_b200
    from pcd_b200 import synthetic
    m = (1 << 17) + 1
    inst = synthetic.make_groth16_instance(ctx, 1, 0, seed=5, num_constraints=m)
    assert pcd_b200.lib.domain_size(1, m + 2) == (200704, 2, 12)
    g = pcd_b200.Groth16(ctx, 1)
    idx = g.index(pcd_b200.ProvingKey(pairing=1, **inst["pk"]),
                  pcd_b200.ConstraintMatrices(1, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"], inst["C"]),
                  precompute=True)
    assert idx.domain_size == 200704
    p = inst["p"]
    r, s = pow(3, 99, p), pow(5, 77, p)
    proof = g.create_proof_with_reduction(idx, inst["z"], codec.int_to_limbs(r), codec.int_to_limbs(s))
    assert np.array_equal(proof.affine_limbs(), synthetic.expected_proof(ctx, inst, r, s, mul=co.fixed_base_mul))
    h = g.witness_map(idx, inst["z"])
    ref_h = co.witness_map(1, inst["A"], inst["B"], inst["C"], m, inst["num_inputs"], inst["z"], threads=16)
    assert np.array_equal(h, ref_h)
    idx.close()


# ---- GM17 ----------------------------------------------------------------------------------
def _gm17_index(ctx, inst, precompute=False):
    import pcd_b200
    pk = pcd_b200.GM17ProvingKey(pairing=inst["pairing"], **inst["pk"])
    cm = pcd_b200.ConstraintMatrices(inst["pairing"], inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"],
                                     inst["C"])
    g = pcd_b200.GM17(ctx, inst["pairing"])
    return g, g.index(pk, cm, precompute)


def test_gm17_golden(ctx):
    for case in codec.load("gm17"):
        pid = case["pairing"]
        inst = dict(pairing=pid, pk=codec.gm17_pk_from_golden(case), num_inputs=case["num_inputs"],
                    num_witness=case["num_witness"], A=codec.csr_from_golden(case["A"]),
                    B=codec.csr_from_golden(case["B"]), C=codec.csr_from_golden(case["C"]))
        z = codec.hex_to_u64(case["z"], 5)
        d1, d2, r = (codec.hex_to_u64(case[k]) for k in ("d1", "d2", "r"))
        for pre in (False, True):
            g, idx = _gm17_index(ctx, inst, pre)
            assert idx.domain_size == case["domain_size"]
            full, h = g.witness_map(idx, z, d1, d2)
            assert codec.u64_to_hex(full) == case["full"]
            assert codec.u64_to_hex(h) == case["h"]
            proof = g.create_proof(idx, z, d1, d2, r)
            assert codec.u64_to_hex(proof.affine_limbs()) == case["proof_affine"]
            assert g.serialize(proof).hex() == case["proof_bytes"]
            idx.close()


@pytest.mark.parametrize("pairing,m,ni", [(0, 1000, 2), (1, 1000, 4), (0, (1 << 12) - 3, 3), (1, (1 << 11) + 5, 2),
                                          (0, 300, 1), (1, 7, 1)])
def test_gm17_vs_oracle_and_trapdoor(ctx, pairing, m, ni):
    inst = synth.make_gm17_instance(pairing, m, seed=700 + m, bitlike=0.4, num_inputs=ni)
    p = codec.FIELD_P[pairing]
    d1, d2, r = pow(3, 201, p), pow(5, 187, p), pow(7, 173, p)
    d1l, d2l, rl = (codec.int_to_limbs(x) for x in (d1, d2, r))
    g, idx = _gm17_index(ctx, inst)
    full, h = g.witness_map(idx, inst["z"], d1l, d2l)
    rfull, rh = co.sap_witness_map(pairing, inst["A"], inst["B"], inst["C"], m, inst["num_inputs"], inst["num_witness"],
                                   inst["z"], d1l, d2l, 8)
    assert np.array_equal(full, rfull)
    assert np.array_equal(h, rh)
    proof = g.create_proof(idx, inst["z"], d1l, d2l, rl)
    ref = co.gm17_prove(pairing, inst["pk"], inst["A"], inst["B"], inst["C"], m, inst["num_inputs"], inst["num_witness"],
                        inst["z"], d1l, d2l, rl, threads=8)
    assert np.array_equal(proof.affine_limbs(), ref)
    assert np.array_equal(proof.affine_limbs(), synth.gm17_trapdoor_proof(inst, d1, d2, r))
    assert g.serialize(proof) == co.serialize_proof(pairing, ref)
    idx.close()
    g2, idx2 = _gm17_index(ctx, inst, precompute=True)
    assert np.array_equal(g2.create_proof(idx2, inst["z"], d1l, d2l, rl).affine_limbs(), ref)
    # the reference's rng contract: prove() draws d1, d2, r in this order
    draws = iter([d1l, d2l, rl])
    assert np.array_equal(g2.prove(idx2, inst["z"], lambda f: next(draws)).affine_limbs(), ref)
    # serialised streams (no lanes) give the same bytes
    ctx.set_concurrency(False)
    try:
        assert np.array_equal(g2.create_proof(idx2, inst["z"], d1l, d2l, rl).affine_limbs(), ref)
    finally:
        ctx.set_concurrency(True)
    idx2.close()


def test_gm17_mixed_radix_sap_domain(ctx):
    """helper side (MNT6-298, Fr = q4): 65540 constraints give a SAP of 131083 rows, beyond q4's 2-adicity -- the domain is
    49 * 2^12 = 200704; the proof must still satisfy GM17's verification equations in the exponent"""
    import pcd_b200
    from pcd_b200 import synthetic
    m = (1 << 16) + 4
    inst = synthetic.make_gm17_instance(ctx, 1, m, seed=8)
    assert inst["domain_size"] == 200704
    g = pcd_b200.GM17(ctx, 1)
    idx = g.index(pcd_b200.GM17ProvingKey(pairing=1, **inst["pk"]),
                  pcd_b200.ConstraintMatrices(1, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"], inst["C"]),
                  precompute=True)
    p = inst["p"]
    d1, d2, r = pow(3, 71, p), pow(5, 61, p), pow(7, 51, p)
    proof = g.create_proof(idx, inst["z"], codec.int_to_limbs(d1), codec.int_to_limbs(d2), codec.int_to_limbs(r))
    assert np.array_equal(proof.affine_limbs(), synthetic.expected_gm17_proof(ctx, inst, d1, d2, r, mul=co.fixed_base_mul))
    # the SAP witness map alone against the C++ oracle
    full, h = g.witness_map(idx, inst["z"], codec.int_to_limbs(d1), codec.int_to_limbs(d2))
    rfull, rh = co.sap_witness_map(1, inst["A"], inst["B"], inst["C"], m, inst["num_inputs"], inst["num_witness"], inst["z"],
                                   codec.int_to_limbs(d1), codec.int_to_limbs(d2), 16)
    assert np.array_equal(full, rfull) and np.array_equal(h, rh)
    idx.close()
