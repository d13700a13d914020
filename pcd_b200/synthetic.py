"""Synthetic inputs for benchmarks, generated without any CPU reference code: random curve points
are k_i * G computed on the GPU (pcdgpu_fixed_base_mul), scalars are seeded numpy draws."""
from __future__ import annotations

import numpy as np

from . import lib as L

R4 = 475922286169261325753349249653048451545124878552823515553267735739164647307408490559963137
Q4 = 475922286169261325753349249653048451545124879242694725395555128576210262817955800483758081
FIELD_P = {L.FIELD_R4: R4, L.FIELD_Q4: Q4}
#: coordinate field of each curve's (base-field) generator encoding
_R = 1 << 320

# Generators (affine, plain integers): MNT4 G1 is arkworks' generator; the others are derived
# deterministically (smallest x, cofactor cleared) -- same values as pcd_b200/csrc/constants.cuh.
_GEN_WORDS = None


def _parse_generators():
    """Read the generator limbs out of csrc/constants.cuh (Montgomery form already)."""
    import os
    import re
    global _GEN_WORDS
    if _GEN_WORDS is not None:
        return _GEN_WORDS
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "constants.cuh")).read()
    out = {}
    for cid, name in ((0, "GenMnt4G1"), (1, "GenMnt4G2"), (2, "GenMnt6G1"), (3, "GenMnt6G2")):
        body = src[src.index("struct %s {" % name):]
        body = body[:body.index("\n};")]
        words = []
        for coord in ("gen_x", "gen_y"):
            m = re.search(coord + r"\(int i\) \{ constexpr u32 v\[\d+\] = \{([^}]*)\}", body)
            words += [int(w.strip().rstrip("u"), 16) for w in m.group(1).split(",")]
        w = np.array(words, dtype=np.uint32)
        out[cid] = w.view(np.uint64).copy()
    _GEN_WORDS = out
    return out


def generator(curve: int) -> np.ndarray:
    """Affine generator of the curve as u64 limbs (Montgomery coordinates)."""
    return _parse_generators()[curve]


def random_limbs(n: int, field: int, seed: int) -> np.ndarray:
    """n uniform field elements as (n, 5) u64 limbs (rejection sampling on 298-bit draws)."""
    p = FIELD_P[field]
    pl = np.array([(p >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(5)], dtype=np.uint64)
    rng = np.random.Generator(np.random.Philox(seed))
    out = np.zeros((n, 5), dtype=np.uint64)
    todo = np.arange(n)
    while len(todo):
        d = rng.integers(0, 2 ** 64, size=(len(todo), 5), dtype=np.uint64)
        d[:, 4] &= np.uint64((1 << (298 - 256)) - 1)
        lt = np.zeros(len(todo), dtype=bool)
        eq = np.ones(len(todo), dtype=bool)
        for i in range(4, -1, -1):
            lt |= eq & (d[:, i] < pl[i])
            eq &= d[:, i] == pl[i]
        out[todo[lt]] = d[lt]
        todo = todo[~lt]
    return out


def random_points_dev(ctx: L.Context, curve: int, n: int, seed: int):
    """n random points of the curve's prime-order group in device memory (torch int64 tensor of
    shape (n, limbs)): k_i * G with k_i uniform, computed by the GPU's fixed-base kernel."""
    import torch
    dev = torch.device("cuda", ctx.device)
    k = torch.from_numpy(random_limbs(n, L.SCALAR_FIELD_OF[0 if curve < 2 else 1], seed).view(np.int64)).to(dev)
    pts = torch.empty((n, L.AFFINE_LIMBS[curve]), dtype=torch.int64, device=dev)
    ctx.fixed_base_mul_dev(curve, generator(curve), k.data_ptr(), n, pts.data_ptr())
    ctx.sync()
    return pts


# ---- synthetic Groth16 instances (benchmarks; no CPU reference code involved) -------------------------
def _limbs_from_ints(vals) -> np.ndarray:
    """list of ints < 2^320 -> (n, 5) u64 limbs"""
    b = b"".join(int(v).to_bytes(40, "little") for v in vals)
    return np.frombuffer(b, dtype="<u8").copy().reshape(-1, 5)


def _omega(field: int, n: int) -> int:
    """generator of the evaluation domain of size n = 7^a 2^b: GENERATOR^((p-1)/n)"""
    p = FIELD_P[field]
    gen = 10 if field == L.FIELD_R4 else 17
    assert (p - 1) % n == 0
    return pow(gen, (p - 1) // n, p)


def _synthetic_rows(rnd, p: int, m: int, bitlike: float):
    """m satisfiable constraints over 2 instance variables (z[0] = 1): rows of (coeff, col) and the assignment"""
    z = [1, rnd.randrange(p)]
    rows_a, rows_b, rows_c = [], [], []
    bit_thr = int(bitlike * 1000)
    for i in range(m):
        nv = len(z)
        if rnd.randrange(1000) < bit_thr:
            z.append(rnd.getrandbits(1))
            rows_a.append(((1, nv),))
            rows_b.append(((1, nv),))
            rows_c.append(((1, nv),))
            continue
        u, v, w = rnd.randrange(nv), rnd.randrange(nv), rnd.randrange(nv)
        sel = rnd.randrange(3)
        ca = 1 if sel == 0 else (p - 1 if sel == 1 else rnd.randrange(1, p))
        ra = ((1, u), (ca, v))
        rb = ((rnd.randrange(1, p), w),) if rnd.randrange(2) else ((1, w), (1, 0))
        a = (z[u] + ca * z[v]) % p
        b = sum(c * z[j] for c, j in rb) % p
        z.append(a * b % p)
        rows_a.append(ra)
        rows_b.append(rb)
        rows_c.append(((1, nv),))
    return z, rows_a, rows_b, rows_c


def _csr(rows, p: int):
    """rows of (coeff, col) -> (row_ptr u32, col u32, val u64[nnz, 5] Montgomery)"""
    R = _R % p
    ptr = np.zeros(len(rows) + 1, dtype=np.uint32)
    cols, vals = [], []
    for i, r in enumerate(rows):
        for c, j in r:
            cols.append(j)
            vals.append(c * R % p)
        ptr[i + 1] = len(cols)
    return ptr, np.array(cols, dtype=np.uint32), _limbs_from_ints(vals)


def _lagrange_at(p: int, omega: int, n: int, tau: int):
    """L_i(tau) on the domain <omega> of size n: Z(tau) w^i / (n (tau - w^i)), one batched inversion"""
    zt = (pow(tau, n, p) - 1) % p
    ninv = pow(n, -1, p)
    ws, dens = [1] * n, [0] * n
    w = 1
    for i in range(n):
        ws[i] = w
        dens[i] = (tau - w) % p
        w = w * omega % p
    pref = [1] * (n + 1)
    for i in range(n):
        pref[i + 1] = pref[i] * dens[i] % p
    inv_all = pow(pref[n], -1, p)
    lag = [0] * n
    cst = zt * ninv % p
    for i in range(n - 1, -1, -1):
        lag[i] = cst * ws[i] % p * (inv_all * pref[i] % p) % p
        inv_all = inv_all * dens[i] % p
    return lag


def make_groth16_instance(ctx: L.Context, pairing: int, log_n: int, seed: int = 20261017, bitlike: float = 0.4,
                          verbose=None, num_constraints: int = None):
    """Satisfiable synthetic R1CS with 2^log_n - 2 constraints and 2 instance variables (domain size
    exactly 2^log_n), its assignment, and a Groth16 proving key with a KNOWN trapdoor.

    Shape: a `bitlike` fraction of the constraints are booleanity checks b*b = b on a
    fresh 0/1 witness (what real circuits are full of); the others are w = <A_i,z> * <B_i,z> with 1-2
    terms per row over earlier variables, coefficients from {1, -1, uniform}.  The key's group
    elements are [scalar]G computed by the GPU's fixed-base kernel from the trapdoor's scalars, so
    every proof can be checked against its discrete logarithms (`expected_logs`).  Returns a dict."""
    import random
    import time
    t0 = time.time()
    field = L.SCALAR_FIELD_OF[pairing]
    p = FIELD_P[field]
    ni = 2
    if num_constraints is None:
        n = 1 << log_n
        m = n - 2
    else:  # any size: the domain is GeneralEvaluationDomain::new(m + ni), possibly mixed radix on q4
        m = num_constraints
        dom = L.domain_size(field, m + ni)
        if dom is None:
            raise ValueError("no evaluation domain for %d constraints on this field" % m)
        n = dom[0]
    rnd = random.Random(seed)
    z, rows_a, rows_b, rows_c = _synthetic_rows(rnd, p, m, bitlike)
    nvars = len(z)
    if verbose:
        verbose("synthetic R1CS: %d constraints, %d variables (%.1f s)" % (m, nvars, time.time() - t0))
    R = _R % p
    csr = lambda rows: _csr(rows, p)
    A, B, C = csr(rows_a), csr(rows_b), csr(rows_c)
    z_mont = _limbs_from_ints([v * R % p for v in z])
    # ---- trapdoor scalars -------------------------------------------------------------------------
    alpha, beta, delta, tau = (rnd.randrange(1, p) for _ in range(4))
    omega = _omega(field, n)
    zt = (pow(tau, n, p) - 1) % p
    lag = _lagrange_at(p, omega, n, tau)
    At, Bt, Ct = [0] * nvars, [0] * nvars, [0] * nvars
    for rows, acc in ((rows_a, At), (rows_b, Bt), (rows_c, Ct)):
        for i, r in enumerate(rows):
            li = lag[i]
            for c, j in r:
                acc[j] += c * li
    for j in range(ni):
        At[j] += lag[m + j]
    At = [x % p for x in At]
    Bt = [x % p for x in Bt]
    Ct = [x % p for x in Ct]
    dinv = pow(delta, -1, p)
    h_sc, cur = [0] * (n - 1), zt * dinv % p
    for i in range(n - 1):
        h_sc[i] = cur
        cur = cur * tau % p
    l_sc = [(beta * At[j] + alpha * Bt[j] + Ct[j]) % p * dinv % p for j in range(ni, nvars)]
    if verbose:
        verbose("trapdoor scalars done (%.1f s)" % (time.time() - t0))
    # ---- group elements on the GPU ----------------------------------------------------------------
    g1, g2 = L.G1_OF[pairing], L.G2_OF[pairing]
    fb = lambda curve, vals: ctx.fixed_base_mul(curve, generator(curve), _limbs_from_ints(vals))
    small1 = fb(g1, [alpha, beta, delta])
    small2 = fb(g2, [beta, delta])
    pk = dict(alpha_g1=small1[0], beta_g1=small1[1], delta_g1=small1[2], beta_g2=small2[0], delta_g2=small2[1],
              a_query=fb(g1, At), b_g1_query=fb(g1, Bt), b_g2_query=fb(g2, Bt), h_query=fb(g1, h_sc),
              l_query=fb(g1, l_sc))
    if verbose:
        verbose("proving key built on the GPU (%.1f s)" % (time.time() - t0))
    az = sum(a * b for a, b in zip(z, At)) % p
    bz = sum(a * b for a, b in zip(z, Bt)) % p
    cz = sum(a * b for a, b in zip(z, Ct)) % p
    l_part = sum(z[ni + j] * l_sc[j] for j in range(nvars - ni)) % p

    def expected_logs(r: int, s: int):
        """discrete logs (base G1 / G2 generator) of the proof (A, B, C) for randomness r, s"""
        a_log = (alpha + az + r * delta) % p
        b_log = (beta + bz + s * delta) % p
        h_part = (az * bz - cz) % p * dinv % p
        c_log = (l_part + h_part + s * a_log + r * b_log - r * s % p * delta) % p
        return a_log, b_log, c_log

    return dict(pairing=pairing, log_n=log_n, m=m, num_inputs=ni, num_witness=nvars - ni, A=A, B=B, C=C, z=z_mont,
                pk=pk, expected_logs=expected_logs, p=p)


def expected_proof(ctx: L.Context, inst, r: int, s: int, mul=None) -> np.ndarray:
    """A || B || C affine limbs computed from the instance's known discrete logs: three scalar multiplications of the
    generators.  `mul(curve, generator_limbs, scalar_limbs) -> points` selects who performs them: the tests and the
    benchmark's gates pass the CPU oracle's double-and-add (an independent implementation); the default is the GPU's
    fixed-base kernel (no CPU code is imported by this package)."""
    a_log, b_log, c_log = inst["expected_logs"](r, s)
    g1, g2 = L.G1_OF[inst["pairing"]], L.G2_OF[inst["pairing"]]
    mul = mul or ctx.fixed_base_mul
    ac = mul(g1, generator(g1), _limbs_from_ints([a_log, c_log]))
    b = mul(g2, generator(g2), _limbs_from_ints([b_log]))
    return np.concatenate([ac[0], b[0], ac[1]])



# ---- synthetic GM17 instances ---------------------------------------------------------------------------
def make_gm17_instance(ctx: L.Context, pairing: int, num_constraints: int, seed: int = 20261017, bitlike: float = 0.4,
                       verbose=None):
    """Satisfiable synthetic R1CS (same shape as make_groth16_instance, 2 instance variables), its assignment and
    an ark-gm17-shaped proving key with a KNOWN trapdoor, built on the GPU.  The SAP has 2m + 3 rows and
    m + 1 extra variables; `expected_logs(d1, d2, r)` gives the discrete logs of the proof every correct prover
    outputs for that randomness (GM17's verification equations in the exponent)."""
    import random
    import time
    t0 = time.time()
    field = L.SCALAR_FIELD_OF[pairing]
    p = FIELD_P[field]
    ni, m = 2, num_constraints
    n = ctx.lib.pcdgpu_sap_domain_size(pairing, m, ni)
    if not n:
        raise ValueError("no evaluation domain for the SAP of %d constraints on this field" % m)
    rnd = random.Random(seed)
    z, rows_a, rows_b, rows_c = _synthetic_rows(rnd, p, m, bitlike)
    nvars = len(z)
    nsap = nvars + m + ni - 1
    A, B, C = _csr(rows_a, p), _csr(rows_b, p), _csr(rows_c, p)
    R = _R % p
    z_mont = _limbs_from_ints([v * R % p for v in z])
    alpha, beta, gamma, tau = (rnd.randrange(1, p) for _ in range(4))
    u = _lagrange_at(p, _omega(field, n), n, tau)
    zt = (pow(tau, n, p) - 1) % p
    ev1, ev2, off = nvars, nvars + m - 1, 2 * m
    At, Ct = [0] * nsap, [0] * nsap
    for i in range(m):
        uadd, usub = u[2 * i] + u[2 * i + 1], u[2 * i] - u[2 * i + 1]
        for c, j in rows_a[i]:
            At[j] += uadd * c
        for c, j in rows_b[i]:
            At[j] += usub * c
        for c, j in rows_c[i]:
            Ct[j] += 4 * u[2 * i] * c
        Ct[ev1 + i] += uadd
    At[0] += u[off]
    Ct[0] += u[off]
    for i in range(1, ni):
        u1, u2 = u[off + 2 * i - 1], u[off + 2 * i]
        At[i] += u1 + u2
        At[0] += u1 - u2
        Ct[i] += 4 * u1
        Ct[ev2 + i] += u1 + u2
    At = [x % p for x in At]
    Ct = [x % p for x in Ct]
    ab, g2_ = (alpha + beta) % p, gamma * gamma % p
    a_sc = [gamma * x % p for x in At]
    c1_sc = [(g2_ * Ct[j] + ab * gamma % p * At[j]) % p for j in range(ni, nsap)]
    c2_sc = [2 * g2_ % p * zt % p * x % p for x in At]
    vk_sc = [(gamma * Ct[j] + ab * At[j]) % p for j in range(ni)]
    gzt_sc, cur = [0] * (n + 1), g2_ * zt % p
    for i in range(n + 1):
        gzt_sc[i] = cur
        cur = cur * tau % p
    if verbose:
        verbose("GM17 trapdoor scalars done: %d constraints, SAP domain %d (%.1f s)" % (m, n, time.time() - t0))
    g1, g2 = L.G1_OF[pairing], L.G2_OF[pairing]
    fb = lambda curve, vals: ctx.fixed_base_mul(curve, generator(curve), _limbs_from_ints(vals))
    small1 = fb(g1, [gamma * zt % p, ab * gamma % p * zt % p, g2_ * zt % p * zt % p])
    pk = dict(a_query=fb(g1, a_sc), b_query=fb(g2, a_sc), c_query_1=fb(g1, c1_sc), c_query_2=fb(g1, c2_sc),
              g_gamma2_z_t=fb(g1, gzt_sc), g_gamma_z=small1[0], h_gamma_z=fb(g2, [gamma * zt % p])[0],
              g_ab_gamma_z=small1[1], g_gamma2_z2=small1[2])
    if verbose:
        verbose("GM17 proving key built on the GPU (%.1f s)" % (time.time() - t0))
    # the SAP assignment's extra variables, for the expected logs
    dot = lambda row: sum(c * z[j] for c, j in row) % p
    full = list(z) + [pow((dot(ra) - dot(rb)) % p, 2, p) for ra, rb in zip(rows_a, rows_b)] + \
        [pow((z[i] - 1) % p, 2, p) for i in range(1, ni)]
    uz = sum(x * y for x, y in zip(full, At)) % p
    psi = sum(z[i] * vk_sc[i] for i in range(ni)) % p

    def expected_logs(d1: int, d2: int, r: int):
        a_log = gamma * ((uz + (r + d1) * zt) % p) % p
        c_log = ((a_log + alpha) * (a_log + beta) - alpha * beta - gamma * psi) % p
        return a_log, a_log, c_log

    return dict(pairing=pairing, m=m, num_inputs=ni, num_witness=nvars - ni, domain_size=n, A=A, B=B, C=C, z=z_mont,
                pk=pk, expected_logs=expected_logs, p=p)


def expected_gm17_proof(ctx: L.Context, inst, d1: int, d2: int, r: int, mul=None) -> np.ndarray:
    """as expected_proof, for GM17 (a proof satisfying the two verification equations in the exponent)"""
    a_log, b_log, c_log = inst["expected_logs"](d1, d2, r)
    g1, g2 = L.G1_OF[inst["pairing"]], L.G2_OF[inst["pairing"]]
    mul = mul or ctx.fixed_base_mul
    ac = mul(g1, generator(g1), _limbs_from_ints([a_log, c_log]))
    b = mul(g2, generator(g2), _limbs_from_ints([b_log]))
    return np.concatenate([ac[0], b[0], ac[1]])
