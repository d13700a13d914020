"""Size-independent check of a GPU Marlin proof for an SRS with a KNOWN trapdoor (beta, gamma), used where the Python
oracle prover is too slow (tests at 2^12+ constraints, bench.py's gate): the AHP verifier's two sumcheck identities are
evaluated from the proof's evaluations with host integers; every commitment must equal [p(beta) + gamma r(beta)] G, the
point computed by the C++ ORACLE's scalar multiplication from a value the GPU obtained by a different kernel (Horner
evaluation, not the MSM); and each batched opening must satisfy KZG10's equation in the exponent.  TEST INFRASTRUCTURE
(imports oracle/)."""
import numpy as np

import c_oracle as co
import codec


def check_in_exponent(snark, ipk, proof, srs_beta, srs_gamma, G_limbs):
    from pcd_b200 import marlin as M
    tr = snark.last_trace
    ops, F = ipk.ops, ipk.ops.F
    p = F.p
    g1 = codec.G1_OF[proof.pairing]
    D = ipk.pc.max_degree
    ch, opening, polys, lcs = tr["challenges"], tr["opening_challenges"], tr["polys"], tr["lcs"]

    def point_of_log(v):
        return co.fixed_base_mul(g1, G_limbs, codec.ints_to_limbs([v % p]), 1)[0]

    def host_eval(limbs, z):
        acc = 0
        for c in reversed([F.dec(x) for x in np.asarray(limbs).reshape(-1, 5)]):
            acc = (acc * z + c) % p
        return acc

    logs = {}
    for rnd in proof.commitments:
        for c in rnd:
            lp = polys[c.label]
            at_beta = ops.evaluate(lp.polynomial, srs_beta)
            log = (at_beta + srs_gamma * (host_eval(lp.rand, srs_beta) if lp.rand is not None else 0)) % p
            assert np.array_equal(c.comm, point_of_log(log)), "commitment to %s" % c.label
            slog = None
            if lp.degree_bound is not None:
                slog = (pow(srs_beta, D - lp.degree_bound, p) * at_beta
                        + srs_gamma * (host_eval(lp.shifted_rand, srs_beta) if lp.shifted_rand is not None else 0)) % p
                assert np.array_equal(c.shifted_comm, point_of_log(slog)), "shifted commitment to %s" % c.label
            logs[c.label] = (log, slog)
    for c in ipk.index_comms:
        logs[c.label] = (ops.evaluate(polys[c.label].polynomial, srs_beta), None)
        assert np.array_equal(c.comm, point_of_log(logs[c.label][0])), "index commitment %s" % c.label
    ev = dict(proof.evaluations)
    # the verifier's view of every LC: evaluate the polynomials behind it at the query point (GPU Horner) -- the two
    # sumcheck LCs must vanish, the others must equal the claimed evaluations
    for label, pl, terms in lcs:
        val = sum(c * (ops.evaluate(polys[l].polynomial, ch[pl]) if l is not None else 1) for c, l in terms) % p
        assert val == (0 if label in M.LC_WITH_ZERO_EVAL else ev[label]), "linear combination %s" % label
    pc = {pl: (w, rv) for pl, w, rv in proof.pc_proof}
    for point_label in sorted({pl for _, pl, _ in lcs}):
        zpt = ch[point_label]
        w, rv = pc[point_label]
        _, _, _, w_poly, rw = tr["combined"][point_label]
        wlog = (ops.evaluate(w_poly, srs_beta) + srs_gamma * sum(c * pow(srs_beta, i, p) for i, c in enumerate(rw))) % p
        assert np.array_equal(w, point_of_log(wlog)), "opening witness at %s" % point_label
        clog = value = counter = 0
        for label, _, terms in sorted((l for l in lcs if l[1] == point_label), key=lambda t: t[0]):
            const = sum(c for c, l in terms if l is None) % p
            lc_value = ((ev[label] if label not in M.LC_WITH_ZERO_EVAL else 0) - const) % p
            named = [(c, l) for c, l in terms if l is not None]
            cj = opening[counter]
            counter += 1
            clog = (clog + cj * sum(c * logs[l][0] for c, l in named)) % p
            value = (value + cj * lc_value) % p
            if len(named) == 1 and logs[named[0][1]][1] is not None:
                cj1 = opening[counter]
                counter += 1
                shift = pow(srs_beta, D - polys[named[0][1]].degree_bound, p)
                clog = (clog + cj1 * (logs[named[0][1]][1] - shift * lc_value)) % p
        assert (clog - value - srs_gamma * (rv or 0)) % p == wlog * (srs_beta - zpt) % p, "KZG equation at %s" % point_label
    return True
