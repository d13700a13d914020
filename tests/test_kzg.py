"""KZG10 commit / open and the dense-polynomial kernels: the oracle against naive definitions and
a known-trapdoor SRS on the CPU; the GPU (through the C ABI / pcd_b200.kzg) against the oracle."""
import numpy as np
import pytest

import c_oracle as co
import codec
import kzg_oracle as ko
import synth


def _mont(vals, field):
    p = codec.FIELD_P[field]
    R = (1 << 320) % p
    return codec.ints_to_limbs([v * R % p for v in vals])


def _unmont(limbs, field):
    p = codec.FIELD_P[field]
    Rinv = pow((1 << 320) % p, -1, p)
    return [codec.limbs_to_int(x) * Rinv % p for x in np.asarray(limbs).reshape(-1, 5)]


def _rand_poly(n, field, seed):
    p = codec.FIELD_P[field]
    rng = np.random.Generator(np.random.Philox(seed))
    return [int.from_bytes(rng.bytes(40), "little") % p for _ in range(n)]


@pytest.mark.parametrize("pairing", [0, 1])
def test_oracle_identities_and_trapdoor(pairing):
    field = pairing
    p = codec.FIELD_P[field]
    g1 = codec.G1_OF[pairing]
    G = synth.generator_limbs(g1)
    beta, gamma = pow(3, 123, p), pow(5, 77, p)
    deg = 40
    pg, pgg = ko.setup(pairing, deg, beta, gamma, G, 2)
    assert pg.shape[0] == deg + 1 and pgg.shape[0] == deg + 2
    for n, nb in ((1, 0), (2, 0), (17, 0), (41, 0), (33, 3), (41, 2)):
        poly = _rand_poly(n, field, 100 + n)
        blind = _rand_poly(nb, field, 200 + n) if nb else None
        z = pow(7, 31 + n, p)
        q, v = ko.poly_divide_linear(p, poly, z)
        assert v == ko.poly_eval(p, poly, z)
        # q (X - z) + p(z) == p
        back = ko.poly_mul_naive(p, q, [(-z) % p, 1]) if q else [0]
        back[0] = (back[0] + v) % p
        assert back[:n] == poly[:n] if n > 1 else back[0] == poly[0]
        c = ko.commit(pairing, pg, pgg, poly, blind)
        log = ko.expected_commit_log(p, beta, gamma, poly, blind)
        assert np.array_equal(c, co.fixed_base_mul(g1, G, codec.ints_to_limbs([log]), 1)[0])
        w, value, rv = ko.open_(pairing, pg, pgg, poly, z, blind)
        assert value == v and (rv is None) == (blind is None)
        if blind:
            assert rv == ko.poly_eval(p, blind, z)
        wlog = ko.expected_open_log(p, beta, gamma, poly, z, blind)
        assert np.array_equal(w, co.fixed_base_mul(g1, G, codec.ints_to_limbs([wlog]), 1)[0])


# ---- GPU ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    import pcd_b200
    c = pcd_b200.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 1000, 65536, 65537, 200001])
def test_gpu_poly_divide_linear(ctx, field, n):
    from pcd_b200 import kzg
    p = codec.FIELD_P[field]
    poly = _rand_poly(n, field, 300 + n)
    z = pow(11, 40 + n % 50, p)
    q, e = kzg.poly_divide_linear(ctx, field, _mont(poly, field), _mont([z], field)[0])
    rq, rv = ko.poly_divide_linear(p, poly, z)
    assert _unmont(e, field)[0] == rv
    assert _unmont(q, field) == rq
    # z = 0 and z = 1 (edge values of the power table)
    for zz in (0, 1):
        q, e = kzg.poly_divide_linear(ctx, field, _mont(poly, field), _mont([zz], field)[0])
        rq, rv = ko.poly_divide_linear(p, poly, zz)
        assert _unmont(e, field)[0] == rv and _unmont(q, field) == rq


@pytest.mark.gpu
@pytest.mark.parametrize("field,na,nb", [(0, 1, 1), (0, 5, 3), (0, 100, 29), (1, 64, 65), (0, 3000, 2000), (1, 70000, 62000)])
def test_gpu_poly_mul(ctx, field, na, nb):
    """DensePolynomial product through the evaluation domain (q4 beyond 2^17 goes through the mixed-radix domain)"""
    from pcd_b200 import kzg
    p = codec.FIELD_P[field]
    a, b = _rand_poly(na, field, 1), _rand_poly(nb, field, 2)
    out = _unmont(kzg.poly_mul(ctx, field, _mont(a, field), _mont(b, field)), field)
    if na * nb <= 10000:
        assert out == ko.poly_mul_naive(p, a, b)
    else:  # evaluate both sides at random points (a size-independent identity)
        for k in range(4):
            x = pow(13, 17 + k, p)
            assert ko.poly_eval(p, out, x) == ko.poly_eval(p, a, x) * ko.poly_eval(p, b, x) % p
        assert len(out) == na + nb - 1


def test_hiding_commitment_draw_count():
    """ark-poly-commit kzg10: Randomness::rand(hiding_bound) draws a polynomial of degree
    calculate_hiding_polynomial_degree(hiding_bound) = hiding_bound + 1, i.e. hiding_bound + 2 field elements; the mirror
    must consume exactly that many draws of the caller's rng (every later draw shifts otherwise)"""
    from pcd_b200 import kzg
    for hb in (0, 1, 2, 7):
        assert kzg.hiding_blinding_coefficients(hb) == hb + 2


@pytest.mark.gpu
@pytest.mark.parametrize("pairing,deg,pre", [(0, 300, False), (1, 300, True), (0, 5000, True)])
def test_gpu_kzg_commit_open(ctx, pairing, deg, pre):
    from pcd_b200 import kzg
    field = pairing
    p = codec.FIELD_P[field]
    g1 = codec.G1_OF[pairing]
    G = synth.generator_limbs(g1)
    beta, gamma = pow(3, 321, p), pow(5, 99, p)
    pg, pgg = ko.setup(pairing, deg, beta, gamma, G, 8)
    powers = kzg.Powers(ctx, pairing, pg, pgg, precompute=pre)
    K = kzg.KZG10(ctx)
    for n, hb in ((deg + 1, None), (deg + 1, 2), (deg // 2, 1), (1, None), (7, 0)):
        poly = _rand_poly(n, field, 500 + n)
        blind_ints = _rand_poly(kzg.hiding_blinding_coefficients(hb), field, 600 + n) if hb is not None else None
        draws = iter(_mont(blind_ints, field)) if blind_ints else None
        c, rand = K.commit(powers, _mont(poly, field), hb, (lambda f: next(draws)) if draws else None)
        if draws is not None:  # every blinding coefficient was consumed, none left over
            assert next(draws, None) is None and rand.blinding_polynomial.shape[0] == hb + 2
        assert np.array_equal(c, ko.commit(pairing, pg, pgg, poly, blind_ints, 8))
        log = ko.expected_commit_log(p, beta, gamma, poly, blind_ints)
        assert np.array_equal(c, co.fixed_base_mul(g1, G, codec.ints_to_limbs([log]), 1)[0])
        z = pow(7, 91 + n, p)
        proof, value = K.open(powers, _mont(poly, field), _mont([z], field)[0], rand)
        w, v, rv = ko.open_(pairing, pg, pgg, poly, z, blind_ints, 8)
        assert np.array_equal(proof.w, w)
        assert _unmont(value, field)[0] == v
        assert (proof.random_v is None) == (rv is None)
        if rv is not None:
            assert _unmont(proof.random_v, field)[0] == rv
        wlog = ko.expected_open_log(p, beta, gamma, poly, z, blind_ints)
        assert np.array_equal(proof.w, co.fixed_base_mul(g1, G, codec.ints_to_limbs([wlog]), 1)[0])
    # the zero polynomial and a polynomial with leading / trailing zero coefficients (upstream skips leading zeros)
    zero = [0] * 9
    c, _ = K.commit(powers, _mont(zero, field))
    assert not c.any() and np.array_equal(c, ko.commit(pairing, pg, pgg, zero, None))
    sparse = [0, 0, 0, 5, 0, pow(3, 50, p), 0, 0]
    c, rand = K.commit(powers, _mont(sparse, field))
    assert np.array_equal(c, ko.commit(pairing, pg, pgg, sparse, None, 2))
    proof, value = K.open(powers, _mont(sparse, field), _mont([0], field)[0], rand)   # opening at z = 0
    w, v, _ = ko.open_(pairing, pg, pgg, sparse, 0, None, 2)
    assert np.array_equal(proof.w, w) and _unmont(value, field)[0] == v == 0
    # a polynomial larger than the committer key is refused (check_degree_is_within_bounds)
    import pcd_b200
    with pytest.raises(ValueError):
        K.commit(powers, _mont(_rand_poly(deg + 2, field, 1), field))
    powers.close()
