"""Proof graphs (include/pcdgpu.h: pcdgpu_set_proof_graphs; an option, off by default): a Groth16 proof replayed from a captured CUDA graph must be
the proof the eager path computes -- checked against the proof's known discrete logarithms (scalar multiplications by
the oracle) for fresh (r, s) on every call, at the three regimes of a PCD step (ECCyclePCD::prove,
/root/reference/src/ec_cycle_pcd/mod.rs:171,179 and data_structures.rs:139-143): the gated large proof (main, 2^18),
the proof whose s g_a + r g1_b runs as two extra MSM lanes (helper, 2^16) and the launch-bound default-circuit proof."""
import numpy as np
import pytest

import c_oracle as co
import codec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import pcd_b200
    return pcd_b200.Context(0)


@pytest.mark.parametrize("pairing,log_n", [(0, 18), (1, 16), (1, 10), (0, 10)])
def test_replayed_proofs_match_discrete_logs(ctx, pairing, log_n):
    import torch

    import pcd_b200
    from pcd_b200 import synthetic
    inst = synthetic.make_groth16_instance(ctx, pairing, log_n, seed=900 + pairing + log_n)
    g = pcd_b200.Groth16(ctx, pairing)
    idx = g.index(pcd_b200.ProvingKey(pairing=pairing, **inst["pk"]),
                  pcd_b200.ConstraintMatrices(pairing, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"],
                                              inst["C"]), precompute=True)
    z = torch.from_numpy(inst["z"].view(np.int64)).to("cuda:0")
    p = inst["p"]
    ctx.set_proof_graphs(True)
    cap0, rep0 = ctx.proof_graph_stats()
    for k in range(5):
        r, s = pow(3, 200 + k, p), pow(11, 90 + k, p)
        proof = g.create_proof_dev(idx, z.data_ptr(), codec.int_to_limbs(r), codec.int_to_limbs(s))
        assert np.array_equal(proof.affine_limbs(), synthetic.expected_proof(ctx, inst, r, s, mul=co.fixed_base_mul)), k
    cap1, rep1 = ctx.proof_graph_stats()
    assert cap1 - cap0 >= 1 and rep1 - rep0 >= 3, "the later calls were expected to replay a captured graph"
    # the host-buffer entry point stages the assignment at a fixed scratch address: same graphs, same proofs
    for k in range(3):
        r, s = pow(5, 300 + k, p), pow(13, 70 + k, p)
        proof = g.create_proof_with_reduction(idx, inst["z"], codec.int_to_limbs(r), codec.int_to_limbs(s))
        assert np.array_equal(proof.affine_limbs(), synthetic.expected_proof(ctx, inst, r, s, mul=co.fixed_base_mul)), k
    # graphs off: eager, same proof
    ctx.set_proof_graphs(False)
    r, s = pow(7, 41, p), pow(2, 99, p)
    proof = g.create_proof_dev(idx, z.data_ptr(), codec.int_to_limbs(r), codec.int_to_limbs(s))
    assert np.array_equal(proof.affine_limbs(), synthetic.expected_proof(ctx, inst, r, s, mul=co.fixed_base_mul))
    idx.close()


def test_context_refuses_a_second_concurrent_prover_call(ctx):
    """include/pcdgpu.h: one context per host thread.  Two threads proving on ONE context: every call either returns the
    right proof or is refused with PCDGPU_E_ARG -- never a proof computed in scratch another call was using"""
    import threading

    import torch

    import pcd_b200
    from pcd_b200 import synthetic
    from pcd_b200.lib import PcdGpuError
    pairing, log_n = 1, 14
    inst = synthetic.make_groth16_instance(ctx, pairing, log_n, seed=4242)
    g = pcd_b200.Groth16(ctx, pairing)
    idx = g.index(pcd_b200.ProvingKey(pairing=pairing, **inst["pk"]),
                  pcd_b200.ConstraintMatrices(pairing, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"],
                                              inst["C"]), precompute=True)
    z = torch.from_numpy(inst["z"].view(np.int64)).to("cuda:0")
    p = inst["p"]
    pairs = [(pow(3, 20 + k, p), pow(7, 30 + k, p)) for k in range(12)]
    expect = [synthetic.expected_proof(ctx, inst, r, s, mul=co.fixed_base_mul) for r, s in pairs]
    g.create_proof_dev(idx, z.data_ptr(), codec.int_to_limbs(pairs[0][0]), codec.int_to_limbs(pairs[0][1]))  # warm
    good, refused, bad = [], [], []

    def work(ks):
        for k in ks:
            r, s = pairs[k]
            try:
                proof = g.create_proof_dev(idx, z.data_ptr(), codec.int_to_limbs(r), codec.int_to_limbs(s))
            except PcdGpuError as e:
                (refused if e.code == -1 else bad).append((k, str(e)))
                continue
            (good if np.array_equal(proof.affine_limbs(), expect[k]) else bad).append((k, "wrong proof"))

    ts = [threading.Thread(target=work, args=(range(i, 12, 2),)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not bad, bad
    assert len(good) + len(refused) == 12 and len(good) >= 1
    idx.close()
