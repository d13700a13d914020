"""GPU parity at the sizes BASELINE.json names and bench.py times: G1 MSM 2^16 .. 2^20 on both curves,
G2 MSM 2^16 on both twists, NTT 2^22 in full and 2^24 through the transform's definition at random output indices,
Groth16 / GM17 proofs at the PCD-step sizes (main 2^18 on MNT4-298, helper 2^16 on MNT6-298) and the tiny
default-circuit proofs (domain 2^9 / 2^10, fresh key, no tables) -- every one against the C++ oracle, bit for bit.

Bases and keys are generated on the GPU (k_i * G by pcdgpu_fixed_base_mul: the oracle's double-and-add would take
minutes at 2^20) but are INPUTS here: what is checked is the MSM / proof over them, computed independently by the
oracle from the same bytes; the bases themselves are validated against the oracle on a random sample and for curve
membership in full."""
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


def gpu_points(ctx, curve, n, seed):
    """n points k_i * G as host limbs + the k_i; a random sample is re-derived by the oracle, all are checked on-curve"""
    from pcd_b200 import synthetic
    pts = synthetic.random_points_dev(ctx, curve, n, seed).cpu().numpy().view(np.uint64)
    k = synthetic.random_limbs(n, codec.SCALAR_FIELD[curve], seed)
    rng = np.random.Generator(np.random.Philox(seed))
    sample = rng.integers(0, n, 64)
    ref = co.fixed_base_mul(curve, synth.generator_limbs(curve), k[sample])
    assert np.array_equal(pts[sample], ref), "GPU-generated bases differ from the oracle's k * G"
    assert co.on_curve(curve, pts, synth.coeff_b_limbs(curve))
    return pts


MSM_CASES = [(0, 16), (0, 18), (0, 20), (2, 16), (2, 18), (2, 20), (1, 16), (3, 16)]


@pytest.mark.parametrize("curve,log_n", MSM_CASES)
def test_msm_baseline_size(ctx, curve, log_n):
    import pcd_b200
    n = 1 << log_n
    pts = gpu_points(ctx, curve, n, 4000 + 10 * curve + log_n)
    pts[5] = 0             # infinity among the bases
    pts[9] = pts[8]        # a duplicate base ...
    for dist in ("U", "W"):
        sc = synth.random_scalars(n, curve, 91 + log_n, dist)
        sc[9] = sc[8]      # ... with the same scalar: equal points meet in every bucket
        sc[10] = codec.int_to_limbs(codec.CURVE_ORDER[curve] - 1)
        ref = co.msm(curve, pts, sc, threads=co.hw_threads())
        assert np.array_equal(ctx.msm(curve, pts, sc), ref), ("variable-base", curve, log_n, dist)
        if dist == "W" or log_n <= 18:
            b = pcd_b200.Bases(ctx, curve, pts, precompute=True)
            try:
                assert np.array_equal(b.msm(sc), ref), ("window tables", curve, log_n, dist)
            finally:
                b.close()


def test_ntt_2_22_vs_oracle(ctx):
    x = codec.random_field_elems(1 << 22, 0, 2222)
    for inv, cos in ((0, 1), (1, 1), (0, 0)):
        assert np.array_equal(ctx.ntt(0, x, inv, cos), co.ntt(0, x, inv, cos, threads=co.hw_threads())), (inv, cos)


def test_ntt_2_24_definition(ctx):
    """the size bench.py times: coset FFT of 2^24 elements checked against out[i] = sum_j x_j (g w^i)^j at 64 random
    indices (n multiplications each in the oracle, independent of any FFT), plus the round trip"""
    n = 1 << 24
    x = np.tile(codec.random_field_elems(1 << 20, 0, 2424), (16, 1))  # the benchmark's input shape
    x[::4097, 0] ^= np.uint64(0x5A5A)                                   # ... made aperiodic
    x[:, 4] &= np.uint64((1 << 40) - 1)
    got = ctx.ntt(0, x, False, True)
    rng = np.random.Generator(np.random.Philox(5))
    idx = np.concatenate([[0, 1, n - 1, n // 2], rng.integers(0, n, 60)]).astype(np.uint64)
    assert np.array_equal(got[idx.astype(np.int64)], co.dft_at(0, x, idx, coset=True))
    assert np.array_equal(ctx.ntt(0, got, True, True), x)


def _groth16(ctx, pairing, log_n, precompute, seed):
    import pcd_b200
    from pcd_b200 import synthetic
    inst = synthetic.make_groth16_instance(ctx, pairing, log_n, seed=seed)
    g = pcd_b200.Groth16(ctx, pairing)
    idx = g.index(pcd_b200.ProvingKey(pairing=pairing, **inst["pk"]),
                  pcd_b200.ConstraintMatrices(pairing, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"],
                                              inst["C"]), precompute=precompute)
    return inst, g, idx


@pytest.mark.parametrize("pairing,log_n", [(0, 18), (1, 16)])
def test_groth16_pcd_step_sizes_vs_oracle(ctx, pairing, log_n):
    """main (MNT4-298, 2^18) and helper (MNT6-298, 2^16) proofs of a PCD step (ECCyclePCD::prove,
    /root/reference/src/ec_cycle_pcd/mod.rs:171,179): the GPU proof over resident window tables == the oracle's prover on
    the same key, matrices, assignment, r and s == the proof's known discrete logarithms"""
    from pcd_b200 import synthetic
    inst, g, idx = _groth16(ctx, pairing, log_n, True, 300 + pairing)
    p = inst["p"]
    r, s = pow(3, 123, p), pow(7, 77, p)
    rl, sl = codec.int_to_limbs(r), codec.int_to_limbs(s)
    proof = g.create_proof_with_reduction(idx, inst["z"], rl, sl)
    ref = co.groth16_prove(pairing, inst["pk"], inst["A"], inst["B"], inst["C"], inst["m"], inst["num_inputs"],
                           inst["num_witness"], inst["z"], rl, sl, threads=co.hw_threads())
    assert np.array_equal(proof.affine_limbs(), ref)
    assert np.array_equal(proof.affine_limbs(), synthetic.expected_proof(ctx, inst, r, s, mul=co.fixed_base_mul))
    assert g.serialize(proof) == co.serialize_proof(pairing, ref)
    h = g.witness_map(idx, inst["z"])
    assert np.array_equal(h, co.witness_map(pairing, inst["A"], inst["B"], inst["C"], inst["m"], inst["num_inputs"],
                                            inst["z"], co.hw_threads()))
    idx.close()


@pytest.mark.parametrize("pairing", [0, 1])
@pytest.mark.parametrize("log_n", [9, 10])
def test_tiny_default_circuit_proofs(ctx, pairing, log_n):
    """the default-circuit proves inside MainCircuit / HelpCircuit synthesis
    (/root/reference/src/ec_cycle_pcd/data_structures.rs:139-143,343-350): domain 2^9 / 2^10, key seen for the first
    time (no window tables), and the same key with tables -- both == the oracle"""
    for precompute in (False, True):
        inst, g, idx = _groth16(ctx, pairing, log_n, precompute, 500 + log_n)
        p = inst["p"]
        for r, s in ((pow(3, 55, p), pow(5, 44, p)), (0, 1), (p - 1, 2)):
            rl, sl = codec.int_to_limbs(r), codec.int_to_limbs(s)
            proof = g.create_proof_with_reduction(idx, inst["z"], rl, sl)
            ref = co.groth16_prove(pairing, inst["pk"], inst["A"], inst["B"], inst["C"], inst["m"], inst["num_inputs"],
                                   inst["num_witness"], inst["z"], rl, sl, threads=4)
            assert np.array_equal(proof.affine_limbs(), ref), (pairing, log_n, precompute, r, s)
        idx.close()


@pytest.mark.parametrize("pairing,log_sap", [(0, 18), (1, 16)])
def test_gm17_pcd_step_sizes_vs_oracle(ctx, pairing, log_sap):
    """GM17 at the same sizes (SAP domain 2^18 on MNT4-298, 2^16 on MNT6-298) against the oracle's prover"""
    import pcd_b200
    from pcd_b200 import synthetic
    m = (1 << (log_sap - 1)) - 2
    inst = synthetic.make_gm17_instance(ctx, pairing, m, seed=4242 + pairing)
    assert inst["domain_size"] == 1 << log_sap
    g = pcd_b200.GM17(ctx, pairing)
    idx = g.index(pcd_b200.GM17ProvingKey(pairing=pairing, **inst["pk"]),
                  pcd_b200.ConstraintMatrices(pairing, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"],
                                              inst["C"]), precompute=True)
    p = inst["p"]
    d1, d2, r = pow(3, 71, p), pow(5, 61, p), pow(7, 51, p)
    d1l, d2l, rl = (codec.int_to_limbs(v) for v in (d1, d2, r))
    proof = g.create_proof(idx, inst["z"], d1l, d2l, rl)
    ref = co.gm17_prove(pairing, inst["pk"], inst["A"], inst["B"], inst["C"], m, inst["num_inputs"], inst["num_witness"],
                        inst["z"], d1l, d2l, rl, threads=co.hw_threads())
    assert np.array_equal(proof.affine_limbs(), ref)
    assert np.array_equal(proof.affine_limbs(), synthetic.expected_gm17_proof(ctx, inst, d1, d2, r, mul=co.fixed_base_mul))
    idx.close()
