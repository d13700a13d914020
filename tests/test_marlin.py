"""Marlin (/root/reference/tests/mnt4_marlin.rs:53-94): the oracle's prover against the AHP verifier
identities and a known-trapdoor SRS on the CPU; the GPU prover (pcd_b200.marlin over the C ABI) against the oracle byte
for byte."""
import numpy as np
import pytest

import c_oracle as co
import codec
import kzg_oracle as ko
import marlin_oracle as mo
import pcd_oracle as o
import synth

FP_OF = {0: o.FR4, 1: o.FQ4}  # scalar field of the pairing
CF_OF = {0: o.FQ4, 1: o.FR4}  # base field of its G1 = the sponge field


class Draws:
    """the caller's rng: one F::rand per call (SplitMix64-based; the real rng stays on the Rust side)"""

    def __init__(self, fp, seed):
        self.fp, self.rng, self.count = fp, o.SplitMix64(seed), 0

    def __call__(self):
        self.count += 1
        return self.rng.field(self.fp.p)


def _coords_of(pairing):
    cf = CF_OF[pairing]
    Rinv = pow(cf.R, -1, cf.p)

    def coords(pt):
        pt = np.asarray(pt, dtype=np.uint64).reshape(2, 5)
        x, y = (codec.limbs_to_int(pt[0]) * Rinv % cf.p, codec.limbs_to_int(pt[1]) * Rinv % cf.p)
        inf = 1 if not pt.any() else 0
        return [x, y, inf]
    return coords


def marlin_setup(pairing, m, seed=5, num_inputs=3):
    fp = FP_OF[pairing]
    r1cs, z = o.synthetic_r1cs(fp, m, num_inputs, seed, 0.3)
    idx = mo.index(r1cs)
    max_degree = max(3 * idx.H.size, 4 * idx.K.size)
    p = fp.p
    beta, gamma = pow(3, 1001, p), pow(5, 777, p)
    g1 = codec.G1_OF[pairing]
    G = synth.generator_limbs(g1)
    pg, pgg = ko.setup(pairing, max_degree, beta, gamma, G, 0)
    return dict(pairing=pairing, fp=fp, r1cs=r1cs, z=z, idx=idx, max_degree=max_degree, beta=beta, gamma=gamma, G=G,
                pg=pg, pgg=pgg, g1=g1)


def oracle_group(S):
    def group(coeffs, shift, blinding):
        assert shift + len(coeffs) <= S["pg"].shape[0]
        parts = []
        if coeffs:
            parts.append(co.msm(S["g1"], S["pg"][shift:shift + len(coeffs)], codec.ints_to_limbs(coeffs), 0))
        if blinding:
            parts.append(co.msm(S["g1"], S["pgg"][:len(blinding)], codec.ints_to_limbs(blinding), 0))
        if not parts:
            return np.zeros(10, dtype=np.uint64)
        return co.point_sum(S["g1"], np.stack(parts))
    return group


def index_comm_coords(S):
    """the index commitments (12 polynomials, no hiding, no bounds), flattened to sponge elements"""
    group, coords = oracle_group(S), _coords_of(S["pairing"])
    comms = [group(c, 0, []) for _, c in S["idx"].index_polys()]
    return comms, [e for c in comms for e in coords(c)]


def oracle_prove(S, seed=99):
    comms, flat = index_comm_coords(S)
    draws = Draws(S["fp"], seed)
    proof, trace = mo.prove(S["idx"], flat, CF_OF[S["pairing"]], S["z"], draws, S["max_degree"], oracle_group(S),
                            _coords_of(S["pairing"]))
    return proof, trace, flat, draws.count


def test_chacha20_block_zero_key():
    """RFC 7539 style known answer: key = 0, counter = 0, nonce = 0 -> first keystream words of ChaCha20"""
    rng = mo.ChaCha20Rng(bytes(32))
    first = b"".join(rng.next_u32().to_bytes(4, "little") for _ in range(8))
    assert first.hex() == "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"


def test_poseidon_sponge_shape():
    for fp in (o.FR4, o.FQ4):
        ark = mo.poseidon_ark(fp)
        assert len(ark) == 39 and all(len(r) == 3 and all(0 <= x < fp.p for x in r) for r in ark)
        s1, s2 = mo.PoseidonSponge(fp), mo.PoseidonSponge(fp)
        s1.absorb([1, 2, 3, 4, 5])
        s2.absorb([1, 2])
        s2.absorb([3, 4, 5])
        assert s1.squeeze(3) == s2.squeeze(3)  # absorbing in pieces is the same stream
        s3 = mo.PoseidonSponge(fp)
        s3.absorb([1, 2, 3, 4, 6])
        assert s3.squeeze(1) != mo_squeeze_again([1, 2, 3, 4, 5], fp)


def mo_squeeze_again(elems, fp):
    s = mo.PoseidonSponge(fp)
    s.absorb(elems)
    return s.squeeze(1)


def test_reindex_is_a_permutation():
    for h, x in ((8, 2), (16, 4), (64, 1), (32, 16)):
        if x == h:
            continue
        if h // x == 1:
            continue
        img = [mo.reindex_by_subdomain(h, x, i) for i in range(h)]
        assert sorted(img) == list(range(h))
        assert img[:x] == [i * (h // x) for i in range(x)]


def test_divide_by_vanishing():
    p = o.R4
    a = [pow(3, i + 1, p) for i in range(37)]
    for n in (4, 8, 16, 64):
        q, r = mo.p_div_vanishing(p, a, n)
        back = mo.p_add(p, mo.p_mul(p, q, [p - 1] + [0] * (n - 1) + [1]), r)
        assert mo.p_trim(back) == a and len(r) <= n


@pytest.mark.parametrize("pairing,m", [(0, 20), (1, 13)])
def test_oracle_marlin_complete(pairing, m):
    """the oracle's proof passes the AHP verifier identities and every KZG check in the exponent; changing an
    evaluation, a challenge-bearing commitment or the witness breaks it"""
    S = marlin_setup(pairing, m)
    assert S["r1cs"].is_satisfied(S["z"])
    proof, trace, flat, ndraws = oracle_prove(S)
    idx, p = S["idx"], S["fp"].p
    # rng draws: 3 blinding coefficients, mask polynomial, hiding polynomials of w, z_a, z_b, mask_poly, g_1 (+ shifted)
    assert ndraws == 3 + 3 * idx.H.size + 4 * 3 + 2 * 3
    coords = _coords_of(pairing)

    def point_of_log(v):
        return co.fixed_base_mul(S["g1"], S["G"], codec.ints_to_limbs([v]), 1)[0]

    args = (idx, flat, CF_OF[pairing], S["z"][:S["r1cs"].num_inputs])
    assert mo.check_proof(*args, proof, trace, S["max_degree"], S["beta"], S["gamma"], point_of_log, coords)
    bad = mo.Proof(proof.commitments, [(l, (v + (l == "t")) % p) for l, v in proof.evaluations], proof.pc_proof)
    assert not mo.check_proof(*args, bad, trace, S["max_degree"], S["beta"], S["gamma"], point_of_log, coords)
    # both sumcheck LCs really evaluate to zero at the challenge points (the AHP verifier's equations)
    ev = dict(proof.evaluations)
    x_at_beta = mo.p_eval(p, o.domain_ifft(idx.X, mo.format_assignment(idx, S["z"])[0]), trace.challenges["beta"])
    for label, pl, terms in mo.linear_combinations(idx, trace.challenges, ev, x_at_beta):
        val = sum(c * (mo.p_eval(p, trace.polys[l].poly, trace.challenges[pl]) if l else 1) for c, l in terms) % p
        assert val == (0 if label in mo.LC_WITH_ZERO_EVAL else ev[label]), label
    # an unsatisfying assignment cannot be proven (the outer sumcheck remainder has a constant term)
    zbad = list(S["z"])
    zbad[-1] = (zbad[-1] + 1) % p
    with pytest.raises(AssertionError):
        mo.prove(idx, flat, CF_OF[pairing], zbad, Draws(S["fp"], 1), S["max_degree"], oracle_group(S), coords)


def test_product_sponge_matches_oracle():
    """pcd_b200/fiat_shamir.py (the product's host-side transcript) against the oracle's restatement"""
    from pcd_b200 import fiat_shamir as fsm
    for field, (f, cf) in enumerate(((o.FR4, o.FQ4), (o.FQ4, o.FR4))):
        a, b = mo.FiatShamirRng(f, cf), fsm.FiatShamirAlgebraicSpongeRng(field, 1 - field)
        assert mo.poseidon_ark(cf) == fsm.PoseidonSponge(1 - field).ark
        for rng in (a, b):
            rng.absorb_bytes(b"MARLIN-2019") if rng is a else rng.absorb_bytes(b"MARLIN-2019")
        a.absorb_native([5, 6, 7]); b.absorb_native_field_elements([5, 6, 7])
        a.absorb_nonnative([f.p - 1, 12345, 1 << 200]); b.absorb_nonnative_field_elements([f.p - 1, 12345, 1 << 200])
        assert a.squeeze_nonnative(4) == b.squeeze_nonnative_field_elements(4)
        a.absorb_native([9]); b.absorb_native_field_elements([9])
        assert a.squeeze_128_bits_nonnative(7) == b.squeeze_128_bits_nonnative_field_elements(7)
        assert a.squeeze_native(3) == b.squeeze_native_field_elements(3)


# ---- GPU ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    import pcd_b200
    c = pcd_b200.Context(0)
    yield c
    c.close()


def gpu_prove(ctx, S, seed=99, precompute=False):
    from pcd_b200 import kzg, marlin
    import pcd_b200
    fp = S["fp"]
    A, B, C = synth.r1cs_to_csr(S["r1cs"])
    cm = pcd_b200.ConstraintMatrices(S["pairing"], S["r1cs"].num_inputs, S["r1cs"].num_witness, A, B, C)
    powers = kzg.Powers(ctx, S["pairing"], S["pg"], S["pgg"], precompute=precompute)
    snark = marlin.MarlinSNARK(ctx, S["pairing"])
    ipk = snark.index(cm, powers, S["max_degree"])
    draws = Draws(fp, seed)
    rng = lambda field: synth.mont_limbs([draws()], fp)[0]
    proof = snark.prove(ipk, synth.mont_limbs(S["z"], fp), rng)
    return proof, ipk, snark, draws.count


def assert_same_proof(gpu, orc):
    for rg, ro in zip(gpu.commitments, orc.commitments):
        assert [c.label for c in rg] == [l for l, _, _ in ro]
        for c, (label, comm, shifted) in zip(rg, ro):
            assert np.array_equal(c.comm, comm), label
            assert (c.shifted_comm is None) == (shifted is None), label
            if shifted is not None:
                assert np.array_equal(c.shifted_comm, shifted), label + " (shifted)"
    assert gpu.evaluations == orc.evaluations
    assert len(gpu.pc_proof) == len(orc.pc_proof)
    for (pl, w, rv), (plo, wo, rvo) in zip(gpu.pc_proof, orc.pc_proof):
        assert pl == plo and np.array_equal(w, wo) and rv == rvo, pl


@pytest.mark.gpu
@pytest.mark.parametrize("pairing,m,pre", [(0, 20, False), (1, 13, False), (0, 300, True), (1, 200, False)])
def test_gpu_marlin_matches_oracle(ctx, pairing, m, pre):
    """the GPU prover's commitments, evaluations and opening proofs equal the oracle's byte for byte; the oracle's proof
    passes the AHP verifier identities and the KZG checks in the exponent (so the GPU's does)"""
    S = marlin_setup(pairing, m)
    proof, trace, flat, ndraws = oracle_prove(S)
    gproof, ipk, snark, gdraws = gpu_prove(ctx, S, precompute=pre)
    assert gdraws == ndraws
    comms, _ = index_comm_coords(S)
    for c, ref in zip(ipk.index_comms, comms):
        assert np.array_equal(c.comm, ref), c.label
    assert snark.last_trace["challenges"] == trace.challenges
    assert snark.last_trace["opening_challenges"] == trace.opening_challenges
    assert_same_proof(gproof, proof)
    coords = _coords_of(pairing)
    point_of_log = lambda v: co.fixed_base_mul(S["g1"], S["G"], codec.ints_to_limbs([v]), 1)[0]
    as_oracle = mo.Proof([[(c.label, c.comm, c.shifted_comm) for c in rnd] for rnd in gproof.commitments],
                         gproof.evaluations, gproof.pc_proof)
    assert mo.check_proof(S["idx"], flat, CF_OF[pairing], S["z"][:S["r1cs"].num_inputs], as_oracle, trace,
                          S["max_degree"], S["beta"], S["gamma"], point_of_log, coords)
    ipk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pairing,m", [(0, 3000), (1, 2500)])
def test_gpu_marlin_larger_trapdoor(ctx, pairing, m):
    """beyond the Python oracle prover's reach: the GPU proof checked in the exponent (tests/marlin_check.py)"""
    import marlin_check
    fp = FP_OF[pairing]
    r1cs, z = o.synthetic_r1cs(fp, m, 2, 17, 0.3)
    nnz = max(sum(len(r) for r in M) for M in (r1cs.A, r1cs.B, r1cs.C))
    h = 1 << (max(r1cs.num_vars, m) - 1).bit_length()
    k = 1 << (nnz - 1).bit_length()
    max_degree = max(3 * h, 4 * k)
    p = fp.p
    beta, gamma = pow(3, 2002, p), pow(5, 555, p)
    G = synth.generator_limbs(codec.G1_OF[pairing])
    pg, pgg = ko.setup(pairing, max_degree, beta, gamma, G, 0)
    S = dict(pairing=pairing, fp=fp, r1cs=r1cs, z=z, pg=pg, pgg=pgg, max_degree=max_degree)
    proof, ipk, snark, ndraws = gpu_prove(ctx, S, seed=3, precompute=True)
    assert ndraws == 3 + 3 * ipk.H.n + 4 * 3 + 2 * 3
    assert marlin_check.check_in_exponent(snark, ipk, proof, beta, gamma, G)
    # a corrupted witness must not survive: the outer sumcheck LC no longer vanishes
    zbad = list(z)
    j = max(i for i, v in enumerate(z) if v > 1)  # not a boolean variable (0 -> 1 would still satisfy b * b = b)
    zbad[j] = (zbad[j] + 1) % p
    assert not r1cs.is_satisfied(zbad)
    S2 = dict(S, z=zbad)
    proof2, ipk2, snark2, _ = gpu_prove(ctx, S2, seed=3)
    with pytest.raises(AssertionError):
        marlin_check.check_in_exponent(snark2, ipk2, proof2, beta, gamma, G)
    ipk.close(); ipk2.close()
