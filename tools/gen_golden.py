#!/usr/bin/env python3
"""Generate tests/golden/*.json from the Python big-int oracle (oracle/pcd_oracle.py).

Run from the repo root:  python tools/gen_golden.py
The reference holds no golden vectors for this path, so these are produced by the
*definitions* in the Python oracle -- naive DFT, double-and-add MSM, the Groth16 prover checked
with a known trapdoor -- and pin the C++ oracle and the CUDA library against them.  All values are
hex strings of the ABI byte encodings (include/pcdgpu.h).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pcd_oracle as o  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FIELDS = {0: o.FR4, 1: o.FQ4}
EXT = {2: (o.FQ2_F, o.FQ4), 3: (o.FQ3_F, o.FR4)}  # ext field -> (field object, base params)
CURVES = {0: (o.MNT4_G1, o.FQ4), 1: (o.MNT4_G2, o.FQ4), 2: (o.MNT6_G1, o.FR4), 3: (o.MNT6_G2, o.FR4)}


def mont_hex(a, fp):
    return o.fp_to_mont_bytes(a, fp).hex()


def felt_hex(F, a, fp):
    """coordinate-field element (Fp / Fp2 / Fp3) -> Montgomery bytes of c0 || c1 || c2"""
    return "".join(mont_hex(c, fp) for c in F.coeffs(a))


def point_hex(curve, fp, P):
    F = curve.F
    if P is None:
        return "00" * (80 * F.degree)
    return felt_hex(F, P[0], fp) + felt_hex(F, P[1], fp)


def scalar_hex(k):
    return o.int_to_repr_bytes(k).hex()


def rand_ext(F, rng, p):
    if F.degree == 1:
        return rng.field(p)
    return tuple(rng.field(p) for _ in range(F.degree))


def gen_fields():
    out = {}
    rng = o.SplitMix64(101)
    for fid, fp in FIELDS.items():
        p = fp.p
        cases = []
        specials = [(0, 0), (1, p - 1), (p - 1, p - 1), (0, 5)]
        for i in range(12):
            a, b = specials[i] if i < len(specials) else (rng.field(p), rng.field(p))
            cases.append({
                "a": mont_hex(a, fp), "b": mont_hex(b, fp),
                "add": mont_hex((a + b) % p, fp), "sub": mont_hex((a - b) % p, fp),
                "mul": mont_hex(a * b % p, fp), "sqr": mont_hex(a * a % p, fp),
                "neg": mont_hex((-a) % p, fp), "inv": mont_hex(pow(a, -1, p) if a else 0, fp),
                "dbl": mont_hex(2 * a % p, fp),
            })
        out[str(fid)] = cases
    for fid, (F, fp) in EXT.items():
        p = fp.p
        cases = []
        for i in range(8):
            a, b = rand_ext(F, rng, p), rand_ext(F, rng, p)
            if i == 0:
                a = F.zero
            cases.append({
                "a": felt_hex(F, a, fp), "b": felt_hex(F, b, fp),
                "add": felt_hex(F, F.add(a, b), fp), "sub": felt_hex(F, F.sub(a, b), fp),
                "mul": felt_hex(F, F.mul(a, b), fp), "sqr": felt_hex(F, F.sqr(a), fp),
                "neg": felt_hex(F, F.neg(a), fp),
                "inv": felt_hex(F, F.inv(a) if not F.is_zero(a) else F.zero, fp),
                "dbl": felt_hex(F, F.add(a, a), fp),
            })
        out[str(fid)] = cases
    return out


def gen_ntt():
    out = []
    rng = o.SplitMix64(202)
    for fid, fp in FIELDS.items():
        p = fp.p
        for log_n in (0, 1, 2, 3, 5, 7):
            n = 1 << log_n
            d = o.domain_new(fp, n)
            assert d.size == n
            vals = [rng.field(p) for _ in range(n)]
            # the definition: naive DFT (+ coset shift / scaling as ark-poly defines them)
            fwd = o.dft_naive(vals, d.omega, p)
            assert fwd == o.domain_fft(d, vals)
            g = fp.generator
            cos = o.dft_naive([v * pow(g, i, p) % p for i, v in enumerate(vals)], d.omega, p)
            assert cos == o.domain_coset_fft(d, vals)
            ninv = pow(n, -1, p)
            inv = [x * ninv % p for x in o.dft_naive(vals, pow(d.omega, -1, p), p)]
            assert inv == o.domain_ifft(d, vals)
            cinv = [x * pow(g, -i, p) % p for i, x in enumerate(inv)]
            assert cinv == o.domain_coset_ifft(d, vals)
            enc = lambda v: "".join(mont_hex(x, fp) for x in v)
            out.append({"field": fid, "log_n": log_n, "input": enc(vals), "fft": enc(fwd), "coset_fft": enc(cos),
                        "ifft": enc(inv), "coset_ifft": enc(cinv)})
    return out


def gen_msm():
    out = []
    rng = o.SplitMix64(303)
    for cid, (curve, fp) in CURVES.items():
        order = curve.order
        g = o.generator(curve)
        for n in (1, 6, 33 if curve.F.degree == 1 else 12):
            pts, sc = [], []
            for i in range(n):
                k = rng.field(order) or 1
                pts.append(curve.mul(g, k))
                sc.append(rng.field(order))
            if n >= 6:
                pts[1] = None            # infinity in the bases
                sc[2] = 0                # zero scalar
                sc[3] = 1                # unit scalar
                sc[4] = order - 1        # -1
                pts[5] = curve.neg(pts[0])  # P and -P ...
                sc[5] = sc[0]            # ... with equal scalars: cancels (exceptional add)
            if n >= 12:
                pts[7] = pts[6]          # duplicate base
                sc[7] = sc[6]            # same scalar: doubling inside a bucket
                sc[8] = 1
                sc[9] = 1
                sc[10] = 2 ** 45
            ref = o.msm_naive(curve, pts, sc)
            assert ref == o.msm_pippenger(curve, pts, sc)
            assert curve.is_on_curve(ref)
            out.append({"curve": cid, "n": n, "bases": "".join(point_hex(curve, fp, P) for P in pts),
                        "scalars": "".join(scalar_hex(k) for k in sc), "result": point_hex(curve, fp, ref)})
        # all-cancelling input: result is the point at infinity
        P = curve.mul(g, 12345)
        out.append({"curve": cid, "n": 2, "bases": point_hex(curve, fp, P) + point_hex(curve, fp, curve.neg(P)),
                    "scalars": scalar_hex(77) * 2, "result": point_hex(curve, fp, None)})
    return out


def csr_of(rows, fp):
    ptr, col, val = [0], [], []
    for r in rows:
        for co, j in r:
            col.append(j)
            val.append(mont_hex(co, fp))
        ptr.append(len(col))
    return {"ptr": ptr, "col": col, "val": "".join(val)}


def gen_groth16():
    out = []
    for pid, pairing in ((0, o.MNT4), (1, o.MNT6)):
        fp = pairing.fr
        p = fp.p
        g1fp = CURVES[0 if pid == 0 else 2][1]
        for (m, bitlike, seed) in ((6, 0.0, 11), (13, 0.4, 12)):
            r1cs, z = o.synthetic_r1cs(fp, m, num_inputs=2, seed=seed, bitlike=bitlike)
            assert r1cs.is_satisfied(z)
            pk = o.groth16_setup(pairing, r1cs, seed=seed + 100)
            rng = o.SplitMix64(seed + 200)
            r, s = rng.field(p), rng.field(p)
            h, d = o.witness_map(r1cs, z)
            assert h[-1] == 0
            proof = o.groth16_prove(pk, r1cs, z, r, s)
            assert proof == o.groth16_prove(pk, r1cs, z, r, s, msm=o.msm_naive)
            assert o.groth16_trapdoor_check(pk, r1cs, z, r, s, proof)
            G1, G2 = pairing.g1, pairing.g2
            ph1 = lambda P: point_hex(G1, g1fp, P)
            ph2 = lambda P: point_hex(G2, g1fp, P)
            out.append({
                "pairing": pid, "m": m, "num_inputs": r1cs.num_inputs, "num_witness": r1cs.num_witness,
                "A": csr_of(r1cs.A, fp), "B": csr_of(r1cs.B, fp), "C": csr_of(r1cs.C, fp),
                "z": "".join(mont_hex(x, fp) for x in z),
                "r": scalar_hex(r), "s": scalar_hex(s),
                "h": "".join(mont_hex(x, fp) for x in h),
                "pk": {
                    "alpha_g1": ph1(pk.alpha_g1), "beta_g1": ph1(pk.beta_g1), "delta_g1": ph1(pk.delta_g1),
                    "beta_g2": ph2(pk.beta_g2), "delta_g2": ph2(pk.delta_g2),
                    "a_query": "".join(ph1(P) for P in pk.a_query),
                    "b_g1_query": "".join(ph1(P) for P in pk.b_g1_query),
                    "b_g2_query": "".join(ph2(P) for P in pk.b_g2_query),
                    "h_query": "".join(ph1(P) for P in pk.h_query),
                    "l_query": "".join(ph1(P) for P in pk.l_query),
                },
                "proof_affine": ph1(proof[0]) + ph2(proof[1]) + ph1(proof[2]),
                "proof_bytes": o.serialize_proof(pairing, proof).hex(),
            })
    return out


def gen_gm17():
    """GM17 proofs (oracle/pcd_oracle.py gm17_*): key with a known trapdoor, the SAP witness map outputs and the
    proof, each checked against GM17's verification equations in the exponent before it is written."""
    out = []
    for pid, pairing in ((0, o.MNT4), (1, o.MNT6)):
        fp = pairing.fr
        p = fp.p
        g1fp = CURVES[0 if pid == 0 else 2][1]
        for (m, ni, bitlike, seed) in ((5, 2, 0.0, 21), (11, 3, 0.4, 22)):
            r1cs, z = o.synthetic_r1cs(fp, m, num_inputs=ni, seed=seed, bitlike=bitlike)
            assert r1cs.is_satisfied(z)
            pk = o.gm17_setup(pairing, r1cs, seed=seed + 100)
            rng = o.SplitMix64(seed + 200)
            d1, d2, r = rng.field(p), rng.field(p), rng.field(p)
            full, h, d = o.sap_witness_map(r1cs, z, d1, d2)
            proof = o.gm17_prove(pk, r1cs, z, d1, d2, r)
            assert proof == o.gm17_prove(pk, r1cs, z, d1, d2, r, msm=o.msm_naive)
            assert o.gm17_trapdoor_check(pk, r1cs, z, d1, d2, r, proof)
            G1, G2 = pairing.g1, pairing.g2
            ph1 = lambda P: point_hex(G1, g1fp, P)
            ph2 = lambda P: point_hex(G2, g1fp, P)
            out.append({
                "pairing": pid, "m": m, "num_inputs": r1cs.num_inputs, "num_witness": r1cs.num_witness,
                "domain_size": d.size,
                "A": csr_of(r1cs.A, fp), "B": csr_of(r1cs.B, fp), "C": csr_of(r1cs.C, fp),
                "z": "".join(mont_hex(x, fp) for x in z),
                "d1": scalar_hex(d1), "d2": scalar_hex(d2), "r": scalar_hex(r),
                "full": "".join(mont_hex(x, fp) for x in full),
                "h": "".join(mont_hex(x, fp) for x in h),
                "pk": {
                    "a_query": "".join(ph1(P) for P in pk.a_query),
                    "b_query": "".join(ph2(P) for P in pk.b_query),
                    "c_query_1": "".join(ph1(P) for P in pk.c_query_1),
                    "c_query_2": "".join(ph1(P) for P in pk.c_query_2),
                    "g_gamma2_z_t": "".join(ph1(P) for P in pk.g_gamma2_z_t),
                    "g_gamma_z": ph1(pk.g_gamma_z), "h_gamma_z": ph2(pk.h_gamma_z),
                    "g_ab_gamma_z": ph1(pk.g_ab_gamma_z), "g_gamma2_z2": ph1(pk.g_gamma2_z2),
                },
                "proof_affine": ph1(proof[0]) + ph2(proof[1]) + ph1(proof[2]),
                "proof_bytes": o.serialize_proof(pairing, proof).hex(),
            })
    return out


def main():
    o.self_check()
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("fields", gen_fields), ("ntt", gen_ntt), ("msm", gen_msm), ("groth16", gen_groth16),
                     ("gm17", gen_gm17)):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        data = fn()
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(data, f, indent=0)
        print("wrote", name)


if __name__ == "__main__":
    main()
