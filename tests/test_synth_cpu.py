"""CPU checks of the synthetic-instance generator and of the C++ oracle's Groth16 prover against
the known-trapdoor discrete logs at a size the Python oracle cannot reach quickly."""
import numpy as np
import pytest

import c_oracle as co
import codec
import synth


@pytest.mark.parametrize("pairing", [0, 1])
def test_oracle_proof_matches_trapdoor(pairing):
    inst = synth.make_instance(pairing, 200, seed=31 + pairing, bitlike=0.3)
    assert inst["r1cs"].is_satisfied(inst["z_int"])
    p = codec.FIELD_P[pairing]
    r, s = 0x1234567 * 3 ** 40 % p, 0x7654321 * 5 ** 50 % p
    proof = co.groth16_prove(pairing, inst["pk"], inst["A"], inst["B"], inst["C"], inst["m"], inst["num_inputs"],
                             inst["num_witness"], inst["z"], codec.int_to_limbs(r), codec.int_to_limbs(s), threads=8)
    assert np.array_equal(proof, synth.trapdoor_proof(inst, r, s))


def test_points_on_curve():
    for cv in (0, 1, 2, 3):
        pts = synth.random_points(20, cv, 3)
        assert co.on_curve(cv, pts, synth.coeff_b_limbs(cv))
