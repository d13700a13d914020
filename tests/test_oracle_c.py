"""Pins the C++ oracle (oracle/c) against the golden vectors produced by the Python big-int oracle
(tools/gen_golden.py).  CPU only."""
import numpy as np
import pytest

import c_oracle as co
import codec

OPS = {"add": 0, "sub": 1, "mul": 2, "sqr": 3, "inv": 4, "neg": 5, "dbl": 10}


@pytest.mark.parametrize("field", [0, 1, 2, 3])
def test_field_ops(field):
    for case in codec.load("fields")[str(field)]:
        a, b = codec.hex_to_u64(case["a"]), codec.hex_to_u64(case["b"])
        for name, op in OPS.items():
            got = co.field_op(field, op, a, b)
            assert codec.u64_to_hex(got) == case[name], (field, name)


def test_ntt_golden():
    for case in codec.load("ntt"):
        x = codec.hex_to_u64(case["input"], 5)
        for name, inv, cos in (("fft", 0, 0), ("coset_fft", 0, 1), ("ifft", 1, 0), ("coset_ifft", 1, 1)):
            for threads in (1, 3):
                got = co.ntt(case["field"], x, inv, cos, threads)
                assert codec.u64_to_hex(got) == case[name], (case["field"], case["log_n"], name)


def test_ntt_roundtrip_and_linearity():
    for field in (0, 1):
        x = codec.random_field_elems(1 << 10, field, 5)
        y = codec.random_field_elems(1 << 10, field, 6)
        for cos in (0, 1):
            fx = co.ntt(field, x, 0, cos, 2)
            assert np.array_equal(co.ntt(field, fx, 1, cos, 2), x)
            fy = co.ntt(field, y, 0, cos, 2)
            s = np.stack([co.field_op(field, 0, x[i], y[i]) for i in range(16)])
            fs = np.stack([co.field_op(field, 0, fx[i], fy[i]) for i in range(len(fx))])
            full = np.stack([co.field_op(field, 0, x[i], y[i]) for i in range(len(x))])
            assert np.array_equal(co.ntt(field, full, 0, cos, 2), fs)
            assert s.shape == (16, 5)


def test_msm_golden():
    for case in codec.load("msm"):
        cv = case["curve"]
        bases = codec.hex_to_u64(case["bases"], codec.POINT_LIMBS[cv])
        sc = codec.hex_to_u64(case["scalars"], 5)
        for threads, c in ((1, 0), (4, 0), (2, 5)):
            got = co.msm(cv, bases, sc, threads, c)
            assert codec.u64_to_hex(got) == case["result"], (cv, case["n"], threads, c)


def test_fixed_base_mul_matches_msm():
    g = codec.load("msm")
    for cv in (0, 1, 2, 3):
        case = [c for c in g if c["curve"] == cv and c["n"] == 6][0]
        bases = codec.hex_to_u64(case["bases"], codec.POINT_LIMBS[cv])
        sc = codec.hex_to_u64(case["scalars"], 5)
        # sum_i k_i P_i computed as sum of single-point multiplications
        parts = np.stack([co.fixed_base_mul(cv, bases[i], sc[i:i + 1], 1)[0] for i in range(6)])
        assert codec.u64_to_hex(co.point_sum(cv, parts)) == case["result"]


def test_groth16_golden():
    for case in codec.load("groth16"):
        pid = case["pairing"]
        g1, g2 = codec.G1_OF[pid], codec.G2_OF[pid]
        A, B, C = (codec.csr_from_golden(case[k]) for k in "ABC")
        z = codec.hex_to_u64(case["z"], 5)
        h = co.witness_map(pid, A, B, C, case["m"], case["num_inputs"], z, threads=2)
        assert codec.u64_to_hex(h) == case["h"]
        pk = {}
        for k, v in case["pk"].items():
            pk[k] = codec.hex_to_u64(v, codec.POINT_LIMBS[g2 if "g2" in k else g1])
        proof = co.groth16_prove(pid, pk, A, B, C, case["m"], case["num_inputs"], case["num_witness"], z,
                                 codec.hex_to_u64(case["r"]), codec.hex_to_u64(case["s"]), threads=2)
        assert codec.u64_to_hex(proof) == case["proof_affine"]
        assert co.serialize_proof(pid, proof).hex() == case["proof_bytes"]
        assert len(case["proof_bytes"]) // 2 == (152 if pid == 0 else 190)


# ---- GM17: C++ oracle vs the Python oracle's golden vectors and the trapdoor check ----------
def test_gm17_golden():
    for case in codec.load("gm17"):
        pid = case["pairing"]
        A, B, C = (codec.csr_from_golden(case[k]) for k in "ABC")
        z = codec.hex_to_u64(case["z"], 5)
        d1, d2, r = (codec.hex_to_u64(case[k]) for k in ("d1", "d2", "r"))
        assert co.sap_domain_size(pid, case["m"], case["num_inputs"]) == case["domain_size"]
        full, h = co.sap_witness_map(pid, A, B, C, case["m"], case["num_inputs"], case["num_witness"], z, d1, d2, 2)
        assert codec.u64_to_hex(full) == case["full"]
        assert codec.u64_to_hex(h) == case["h"]
        pk = codec.gm17_pk_from_golden(case)
        proof = co.gm17_prove(pid, pk, A, B, C, case["m"], case["num_inputs"], case["num_witness"], z, d1, d2, r, 2)
        assert codec.u64_to_hex(proof) == case["proof_affine"]
        assert co.serialize_proof(pid, proof).hex() == case["proof_bytes"]


@pytest.mark.parametrize("pairing", [0, 1])
def test_gm17_trapdoor(pairing):
    """a 300-constraint GM17 proof from the C++ oracle satisfies the verification equations in the exponent"""
    import synth
    inst = synth.make_gm17_instance(pairing, 300, seed=41 + pairing, bitlike=0.4, num_inputs=3)
    p = codec.FIELD_P[pairing]
    d1, d2, r = pow(3, 150, p), pow(5, 140, p), pow(7, 130, p)
    proof = co.gm17_prove(pairing, inst["pk"], inst["A"], inst["B"], inst["C"], 300, inst["num_inputs"],
                          inst["num_witness"], inst["z"], codec.int_to_limbs(d1), codec.int_to_limbs(d2),
                          codec.int_to_limbs(r), threads=4)
    assert np.array_equal(proof, synth.gm17_trapdoor_proof(inst, d1, d2, r))
