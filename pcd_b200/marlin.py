"""Host-side mirror of ark-marlin's prover for the reference's Marlin configuration
(/root/reference/tests/mnt4_marlin.rs:72-75: `MainSNARK = MarlinSNARK<Fr, Fq, MarlinKZG10<MNT4_298, DensePolynomial<Fr>>,
FS4, TestMarlinConfig>`, `HelpSNARK` the same over MNT6-298; `TestMarlinConfig::FOR_RECURSION = true`, :62-66), reached
through `IC::MainSNARK::prove` / `IC::HelpSNARK::prove` (/root/reference/src/ec_cycle_pcd/mod.rs:171,179).

Names and argument meaning follow ark-marlin (branch `constraints`): `AHPForR1CS::{index, prover_first_round,
prover_second_round, prover_third_round, construct_linear_combinations}`, `Marlin::prove`, and ark-poly-commit's
`MarlinKZG10::{commit, open_combinations}`.  Every vector operation -- FFTs over H, K and the product domains, sparse
matrix products, pointwise arithmetic, batch inversion, divisions by vanishing / linear polynomials, polynomial
evaluation and all commitments (MSMs over the SRS) -- runs in libpcdgpu.so on device-resident vectors; the host side
holds only challenge-sized data: the Fiat-Shamir sponge (pcd_b200/fiat_shamir.py), the verifier challenges and the
scalars of the linear combinations.  There is no CPU path: without the library / a B200 nothing here computes.

The caller's rng stays on its side of the boundary (as r, s do for Groth16): `rng(field)` returns one `F::rand` as
Montgomery limbs; the draw ORDER is ark-marlin's (w, z_a, z_b blinding, mask polynomial, then the commitments' hiding
polynomials round by round).  ark-marlin is un-vendored and un-pinned in the reference, so this mirror is checked
against oracle/marlin_oracle.py (a from-structure restatement) and, through it, against the AHP verifier identities and
a known-trapdoor SRS -- parity with arkworks' bytes is unpinned."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field as dc_field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import kzg
from . import lib as L
from .fiat_shamir import FiatShamirAlgebraicSpongeRng
from .snark import ConstraintMatrices
from .synthetic import FIELD_P

PROTOCOL_NAME = b"MARLIN-2019"
ZK_BOUND = 1
BASE_FIELD_OF = {L.MNT4_298: L.FIELD_Q4, L.MNT6_298: L.FIELD_R4}  # G1 coordinates = the sponge's native field
_GEN = {L.FIELD_R4: 10, L.FIELD_Q4: 17}
ADD, SUB, MUL, RSUB = 0, 1, 2, 3
NONE = 0xFFFFFFFF


class _Fld:
    """conversions between Python integers and the ABI's Montgomery limbs (challenge-sized data only)"""

    def __init__(self, field: int):
        self.field, self.p = field, FIELD_P[field]
        self.R = (1 << 320) % self.p
        self.Rinv = pow(self.R, -1, self.p)

    def enc(self, v: int) -> np.ndarray:
        return np.frombuffer(((v % self.p) * self.R % self.p).to_bytes(40, "little"), dtype="<u8").copy()

    def enc_many(self, vals: Sequence[int]) -> np.ndarray:
        b = b"".join(((v % self.p) * self.R % self.p).to_bytes(40, "little") for v in vals)
        return np.frombuffer(b, dtype="<u8").copy().reshape(-1, 5)

    def dec(self, limbs) -> int:
        return int.from_bytes(np.ascontiguousarray(limbs, dtype="<u8").tobytes(), "little") * self.Rinv % self.p


def _vp(x):
    return ctypes.c_void_p(x)


class DVec:
    """a vector of field elements in device memory (pcdgpu_dev_alloc); `view` gives non-owning windows"""

    def __init__(self, ctx: L.Context, field: int, n: int, ptr: Optional[int] = None):
        self.ctx, self.field, self.n = ctx, field, n
        self._own = ptr is None
        if ptr is None:
            out = ctypes.c_void_p()
            ctx._check(ctx.lib.pcdgpu_dev_alloc(ctx.h, max(n, 1) * 40, ctypes.byref(out)))
            ptr = out.value
        self.ptr = ptr

    @classmethod
    def from_host(cls, ctx, field, limbs) -> "DVec":
        a = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 5)
        v = cls(ctx, field, a.shape[0])
        if a.shape[0]:
            ctx._check(ctx.lib.pcdgpu_dev_upload(ctx.h, _vp(v.ptr), _vp(a.ctypes.data), a.nbytes))
        return v

    @classmethod
    def zeros(cls, ctx, field, n) -> "DVec":
        v = cls(ctx, field, n)
        ctx._check(ctx.lib.pcdgpu_dev_zero(ctx.h, _vp(v.ptr), n * 40))
        return v

    def host(self) -> np.ndarray:
        out = np.zeros((self.n, 5), dtype=np.uint64)
        if self.n:
            self.ctx._check(self.ctx.lib.pcdgpu_dev_download(self.ctx.h, _vp(out.ctypes.data), _vp(self.ptr), out.nbytes))
        return out

    def view(self, lo: int, hi: Optional[int] = None) -> "DVec":
        hi = self.n if hi is None else hi
        assert 0 <= lo <= hi <= self.n
        v = DVec(self.ctx, self.field, hi - lo, self.ptr + lo * 40)
        v._keep = self  # the window keeps its buffer alive
        return v

    def resized(self, n: int) -> "DVec":
        """a copy of length n (truncated, or padded with zeros)"""
        out = DVec(self.ctx, self.field, n)
        k = min(n, self.n)
        lib, h = self.ctx.lib, self.ctx.h
        if k:
            self.ctx._check(lib.pcdgpu_dev_copy(h, _vp(out.ptr), _vp(self.ptr), k * 40))
        if n > k:
            self.ctx._check(lib.pcdgpu_dev_zero(h, _vp(out.ptr + k * 40), (n - k) * 40))
        return out

    def free(self):
        if self._own and self.ptr and self.ctx.h:
            self.ctx.lib.pcdgpu_dev_free(self.ctx.h, _vp(self.ptr))
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Domain:
    """ark-poly GeneralEvaluationDomain::new(min_size) over the field (radix 2, or 7^a 2^b on q4)"""

    def __init__(self, field: int, min_size: int):
        shape = L.domain_size(field, max(min_size, 1))
        if shape is None:
            raise ValueError("no evaluation domain of %d elements over field %d" % (min_size, field))
        self.field = field
        self.n, self.pow7, self.pow2 = shape
        p = FIELD_P[field]
        self.omega = pow(_GEN[field], (p - 1) // self.n, p)

    def vanishing_at(self, x: int) -> int:
        p = FIELD_P[self.field]
        return (pow(x, self.n, p) - 1) % p


class _Ops:
    """DensePolynomial / EvaluationsOnDomain arithmetic on device vectors (thin calls into the C ABI)"""

    def __init__(self, ctx: L.Context, field: int):
        self.ctx, self.field, self.F = ctx, field, _Fld(field)
        self.lib, self.h = ctx.lib, ctx.h

    def _chk(self, rc):
        self.ctx._check(rc)

    def binary(self, op, out: DVec, a: DVec, b: DVec, n: int):
        self._chk(self.lib.pcdgpu_vec_binary_dev(self.h, self.field, op, _vp(out.ptr), _vp(a.ptr), _vp(b.ptr), n))

    def scalar(self, op, out: DVec, a: DVec, s: int, n: Optional[int] = None):
        sl = self.F.enc(s)
        self._chk(self.lib.pcdgpu_vec_scalar_dev(self.h, self.field, op, _vp(out.ptr), _vp(a.ptr), _vp(sl.ctypes.data),
                                                 a.n if n is None else n))

    def axpy(self, y: DVec, s: int, x: DVec):
        """y[:len(x)] += s x"""
        assert y.n >= x.n
        sl = self.F.enc(s)
        self._chk(self.lib.pcdgpu_vec_axpy_dev(self.h, self.field, _vp(y.ptr), _vp(sl.ctypes.data), _vp(x.ptr), x.n))

    def add(self, a: DVec, b: DVec) -> DVec:
        if a.n < b.n:
            a, b = b, a
        out = a.resized(a.n)
        self.binary(ADD, out, out, b, b.n)
        return out

    def sub(self, a: DVec, b: DVec) -> DVec:
        out = a.resized(max(a.n, b.n))
        self.binary(SUB, out, out, b, b.n)
        return out

    def scaled(self, a: DVec, s: int) -> DVec:
        out = DVec(self.ctx, self.field, a.n)
        self.scalar(MUL, out, a, s)
        return out

    def mul_inplace(self, a: DVec, b: DVec):
        assert a.n == b.n
        self.binary(MUL, a, a, b, a.n)

    def inverse_inplace(self, a: DVec):
        self._chk(self.lib.pcdgpu_vec_inverse_dev(self.h, self.field, _vp(a.ptr), a.n))

    def powers(self, base: int, scale: int, n: int) -> DVec:
        out = DVec(self.ctx, self.field, n)
        b, s = self.F.enc(base), self.F.enc(scale)
        self._chk(self.lib.pcdgpu_vec_powers_dev(self.h, self.field, _vp(out.ptr), _vp(b.ctypes.data), _vp(s.ctypes.data), n))
        return out

    def gather(self, src: DVec, d_index: int, n: int) -> DVec:
        out = DVec(self.ctx, self.field, n)
        self._chk(self.lib.pcdgpu_vec_gather_dev(self.h, self.field, _vp(out.ptr), _vp(src.ptr), _vp(d_index), n))
        return out

    def ntt(self, v: DVec, dom: Domain, inverse: bool):
        assert v.n == dom.n
        self._chk(self.lib.pcdgpu_ntt_general_dev(self.h, self.field, _vp(v.ptr), dom.pow7, dom.pow2, int(inverse), 0))

    def fft(self, dom: Domain, coeffs: DVec) -> DVec:
        assert coeffs.n <= dom.n
        v = coeffs.resized(dom.n)
        self.ntt(v, dom, False)
        return v

    def ifft(self, dom: Domain, evals: DVec) -> DVec:
        v = evals.resized(dom.n)
        self.ntt(v, dom, True)
        return v

    def poly_mul(self, a: DVec, b: DVec) -> DVec:
        n = a.n + b.n - 1
        dom = Domain(self.field, n)
        fa, fb = self.fft(dom, a), self.fft(dom, b)
        self.mul_inplace(fa, fb)
        self.ntt(fa, dom, True)
        fb.free()
        fa.n = n  # the product's coefficients; the buffer stays dom.n long
        return fa

    def evaluate(self, a: DVec, z: int) -> int:
        out = np.zeros(5, dtype=np.uint64)
        zl = self.F.enc(z)
        self._chk(self.lib.pcdgpu_poly_eval_dev(self.h, self.field, _vp(a.ptr), a.n, _vp(zl.ctypes.data), _vp(out.ctypes.data)))
        return self.F.dec(out)

    def divide_by_vanishing(self, a: DVec, N: int) -> Tuple[DVec, DVec]:
        q = DVec(self.ctx, self.field, max(a.n - N, 0))
        r = DVec(self.ctx, self.field, N)
        self._chk(self.lib.pcdgpu_poly_divide_vanishing_dev(self.h, self.field, _vp(a.ptr), a.n, N, _vp(q.ptr), _vp(r.ptr)))
        return q, r

    def divide_by_linear(self, a: DVec, z: int) -> DVec:
        q = DVec(self.ctx, self.field, max(a.n - 1, 0))
        zl = self.F.enc(z)
        self._chk(self.lib.pcdgpu_poly_divide_linear_dev(self.h, self.field, _vp(a.ptr), a.n, _vp(zl.ctypes.data), _vp(q.ptr), None))
        return q

    def matvec(self, csr, x: DVec, m: int) -> DVec:
        out = DVec(self.ctx, self.field, m)
        self._chk(self.lib.pcdgpu_csr_matvec_dev(self.h, csr, _vp(x.ptr), _vp(out.ptr)))
        return out

    def upload_u32(self, a: np.ndarray) -> DVec:
        a = np.ascontiguousarray(a, dtype=np.uint32)
        v = DVec(self.ctx, self.field, (a.nbytes + 39) // 40)
        self._chk(self.lib.pcdgpu_dev_upload(self.h, _vp(v.ptr), _vp(a.ctypes.data), a.nbytes))
        return v


def draw_many(rng, field: int, n: int) -> np.ndarray:
    """n consecutive `F::rand` draws of the caller's rng as (n, 5) Montgomery limbs.  An rng object may offer
    `many(field, n)` for bulk draws (the mask polynomial is 3|H| of them); the order is the same either way."""
    if hasattr(rng, "many"):
        out = np.ascontiguousarray(rng.many(field, n), dtype=np.uint64).reshape(-1, 5)
        if out.shape[0] != n:
            raise ValueError("rng.many returned %d draws, %d were asked for" % (out.shape[0], n))
        return out
    return np.stack([rng(field) for _ in range(n)]).astype(np.uint64)


# ---- ark-poly-commit marlin_pc ------------------------------------------------------------------------------------
@dataclass
class LabeledPolynomial:
    label: str
    polynomial: DVec
    degree_bound: Optional[int] = None
    hiding_bound: Optional[int] = None
    rand: Optional[np.ndarray] = None  # blinding polynomial (host limbs; hiding_bound + 2 coefficients)
    shifted_rand: Optional[np.ndarray] = None


@dataclass
class Commitment:
    """marlin_pc::Commitment { comm, shifted_comm }: affine G1 limbs"""
    label: str
    comm: np.ndarray
    shifted_comm: Optional[np.ndarray] = None


@dataclass
class Proof:
    """ark-marlin `Proof`: commitments per round, evaluations (sorted by label), the batched opening proof
    (`BatchLCProof`: one kzg10::Proof { w, random_v } per query point)"""
    pairing: int
    commitments: List[List[Commitment]]
    evaluations: List[Tuple[str, int]]
    pc_proof: List[Tuple[str, np.ndarray, Optional[int]]]


class MarlinKZG10:
    """the committer-key half of ark-poly-commit's MarlinKZG10 over resident powers (kzg.Powers)"""

    def __init__(self, ctx: L.Context, powers: kzg.Powers, max_degree: int):
        self.ctx, self.powers, self.max_degree = ctx, powers, max_degree
        if powers.powers_of_g.n < max_degree + 1:
            raise ValueError("the SRS holds %d powers, max_degree %d needs %d" % (powers.powers_of_g.n, max_degree, max_degree + 1))

    def _msm(self, coeffs: DVec, shift: int, rand: Optional[np.ndarray]) -> np.ndarray:
        out = np.zeros(L.AFFINE_LIMBS[self.powers.curve], dtype=np.uint64)
        n_rand = 0 if rand is None else rand.shape[0]
        d_rand = DVec.from_host(self.ctx, coeffs.field, rand) if n_rand else None
        self.ctx._check(self.ctx.lib.pcdgpu_kzg_commit_dev(
            self.ctx.h, self.powers.powers_of_g.h, shift, _vp(coeffs.ptr), coeffs.n, self.powers.powers_of_gamma_g.h,
            _vp(d_rand.ptr) if d_rand else None, n_rand, _vp(out.ctypes.data)))
        return out

    def commit(self, polys: Sequence[LabeledPolynomial], rng: Callable[[int], np.ndarray]) -> List[Commitment]:
        """MarlinKZG10::commit: per polynomial the hiding polynomial is drawn (hiding_bound + 2 coefficients), then the
        one of the shifted commitment when a degree bound is enforced; shifted_comm = commit over
        powers_of_g[max_degree - bound ..]"""
        out = []
        for lp in polys:
            field = lp.polynomial.field
            if lp.hiding_bound is not None:
                k = kzg.hiding_blinding_coefficients(lp.hiding_bound)
                lp.rand = draw_many(rng, field, k)
                if lp.degree_bound is not None:
                    lp.shifted_rand = draw_many(rng, field, k)
            c = Commitment(lp.label, self._msm(lp.polynomial, 0, lp.rand))
            if lp.degree_bound is not None:
                if lp.polynomial.n > lp.degree_bound + 1:
                    raise ValueError("%s exceeds its degree bound" % lp.label)
                c.shifted_comm = self._msm(lp.polynomial, self.max_degree - lp.degree_bound, lp.shifted_rand)
            out.append(c)
        return out


# ---- ark-marlin ahp::indexer -------------------------------------------------------------------------------------
def _next_pow2(n: int) -> int:
    return 1 << max(n - 1, 0).bit_length()


def _reindex(h: int, x: int, idx: np.ndarray) -> np.ndarray:
    """EvaluationDomain::reindex_by_subdomain, vectorised"""
    period = h // x
    i = idx.astype(np.int64) - x
    return np.where(idx < x, idx.astype(np.int64) * period, i + i // max(period - 1, 1) + 1)


@dataclass
class MatrixArithmetization:
    row: DVec
    col: DVec
    val: DVec
    row_col: DVec
    evals_on_K_row: DVec
    evals_on_K_col: DVec
    evals_on_K_val: DVec


class IndexProverKey:
    """ark-marlin `IndexProverKey`: the index (matrices and their arithmetization, resident on the GPU), the index
    commitments and the committer key"""

    def __init__(self, ctx: L.Context, cm: ConstraintMatrices, pc: MarlinKZG10):
        self.ctx, self.pairing, self.pc = ctx, cm.pairing, pc
        field = L.SCALAR_FIELD_OF[cm.pairing]
        self.field, self.ops = field, _Ops(ctx, field)
        ops, F = self.ops, self.ops.F
        self.num_inputs_orig = cm.num_instance_variables
        self.num_inputs = _next_pow2(cm.num_instance_variables)  # pad_input_for_indexer_and_prover
        shift = self.num_inputs - self.num_inputs_orig
        self.num_witness = cm.num_witness_variables
        nv = self.num_inputs + self.num_witness
        nc = cm.num_constraints
        self.n = n = max(nv, nc)  # make_matrices_square
        mats = []
        for ptr, col, val in (cm.a, cm.b, cm.c):
            ptr = np.asarray(ptr, dtype=np.int64)
            col = np.asarray(col, dtype=np.int64)
            val = np.ascontiguousarray(val, dtype=np.uint64).reshape(-1, 5)
            rows = np.repeat(np.arange(nc, dtype=np.int64), np.diff(ptr))
            col = np.where(col < self.num_inputs_orig, col, col + shift)
            order = np.lexsort((col, rows))  # row major, columns ascending inside a row
            mats.append((rows[order], col[order], val[order]))
        self.num_non_zero = max(len(m[0]) for m in mats)
        self.H, self.K, self.X = Domain(field, n), Domain(field, self.num_non_zero), Domain(field, self.num_inputs)
        h, k, x = self.H.n, self.K.n, self.X.n
        if h % x or h == x:
            raise ValueError("the input domain must be a proper subgroup of H")
        if pc.max_degree < max(3 * h, 4 * k):
            raise ValueError("max_degree %d is too small for |H| = %d, |K| = %d" % (pc.max_degree, h, k))
        self.h_elems = ops.powers(self.H.omega, 1, h)
        self._keep = []
        # the matrices as CSR for z_A, z_B (rows padded to n), their transposes re-indexed onto H for t(X)
        self.csr, self.csr_t = [], []
        for rows, cols, vals in mats:
            self.csr.append(self._upload_csr(rows, cols, vals, n, n))
            vi = _reindex(h, x, cols)
            o = np.argsort(vi, kind="stable")
            self.csr_t.append(self._upload_csr(vi[o], rows[o], vals[o], h, h))
        # w_evals gather maps (prover_first_round): H position -> index into the formatted assignment / into x_evals
        pos = np.arange(h, dtype=np.int64)
        ratio = h // x
        on_x = pos % ratio == 0
        self.w_index = ops.upload_u32(np.where(on_x, NONE, x + pos - pos // ratio - 1))
        self.x_index = ops.upload_u32(np.where(on_x, NONE, pos))
        # arithmetize_matrix ("the transpose of M"): row(k) = the VARIABLE's element of H, col(k) = the CONSTRAINT's,
        # val(k) = M_k / u_H(row, row) = M_k row(k) / |H|   (u_H(x, x) = |H| x^(|H|-1) and x^|H| = 1 on H)
        self.arith: List[MatrixArithmetization] = []
        h_inv = pow(h, -1, F.p)
        for rows, cols, vals in mats:
            cnt = len(rows)
            vi = _reindex(h, x, cols)
            pad = k - cnt
            ri = np.concatenate([vi, np.full(pad, vi[-1])])
            ci = np.concatenate([rows, np.full(pad, rows[-1])])
            d_ri, d_ci = ops.upload_u32(ri), ops.upload_u32(ci)
            ev_row = ops.gather(self.h_elems, d_ri.ptr, k)
            ev_col = ops.gather(self.h_elems, d_ci.ptr, k)
            ev_val = DVec.from_host(ctx, field, np.concatenate([vals, np.zeros((pad, 5), dtype=np.uint64)]))
            ops.mul_inplace(ev_val, ev_row)
            ops.scalar(MUL, ev_val, ev_val, h_inv)
            rc = DVec(ctx, field, k)
            ops.binary(MUL, rc, ev_row, ev_col, k)
            self.arith.append(MatrixArithmetization(ops.ifft(self.K, ev_row), ops.ifft(self.K, ev_col),
                                                    ops.ifft(self.K, ev_val), ops.ifft(self.K, rc), ev_row, ev_col, ev_val))
        # index commitments (no hiding, no degree bounds) and the hash of the verifier key that seeds the transcript
        self.index_comms = [Commitment(label, pc._msm(poly, 0, None)) for label, poly in self.index_polys()]
        fs = FiatShamirAlgebraicSpongeRng(field, BASE_FIELD_OF[self.pairing])
        fs.absorb_native_field_elements(e for c in self.index_comms for e in point_to_field_elements(self.pairing, c.comm))
        self.vk_hash = fs.squeeze_native_field_elements(1)[0]

    def index_polys(self) -> List[Tuple[str, DVec]]:
        out = []
        for name, m in zip("abc", self.arith):
            out += [(name + "_row", m.row), (name + "_col", m.col), (name + "_val", m.val), (name + "_row_col", m.row_col)]
        return out

    def _upload_csr(self, rows, cols, vals, m, ncols):
        ptr = np.zeros(m + 1, dtype=np.uint32)
        ptr[1:] = np.cumsum(np.bincount(rows, minlength=m))
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        vals = np.ascontiguousarray(vals, dtype=np.uint64)
        out = ctypes.c_void_p()
        self.ctx._check(self.ctx.lib.pcdgpu_csr_upload(self.ctx.h, self.field, m, ncols, _vp(ptr.ctypes.data),
                                                       _vp(cols.ctypes.data), _vp(vals.ctypes.data), ctypes.byref(out)))
        return out

    def close(self):
        for c in self.csr + self.csr_t:
            self.ctx.lib.pcdgpu_csr_free(c)
        self.csr, self.csr_t = [], []


def point_to_field_elements(pairing: int, pt: np.ndarray) -> List[int]:
    """`ToConstraintField` of a short-Weierstrass affine point: x, y, infinity"""
    F = _Fld(BASE_FIELD_OF[pairing])
    pt = np.asarray(pt, dtype=np.uint64).reshape(2, 5)
    return [F.dec(pt[0]), F.dec(pt[1]), 0 if pt.any() else 1]


# ---- ark-marlin ahp::prover + lib.rs ------------------------------------------------------------------------------
LC_WITH_ZERO_EVAL = ("inner_sumcheck", "outer_sumcheck")
LinearCombination = Tuple[str, str, List[Tuple[int, Optional[str]]]]  # label, query point label, (coeff, poly | None)


class AHPForR1CS:
    @staticmethod
    def construct_linear_combinations(pk: IndexProverKey, ch: Dict[str, int], ev: Dict[str, int],
                                      x_at_beta: int) -> List[LinearCombination]:
        p = FIELD_P[pk.field]
        al, be, ga = ch["alpha"], ch["beta"], ch["gamma"]
        eta = (ch["eta_a"], ch["eta_b"], ch["eta_c"])
        vha, vhb = pk.H.vanishing_at(al), pk.H.vanishing_at(be)
        vxb, vkg = pk.X.vanishing_at(be), pk.K.vanishing_at(ga)
        r_alpha_at_beta = (vha - vhb) * pow(al - be, -1, p) % p
        z_b, t, g_1, g_2 = ev["z_b"], ev["t"], ev["g_1"], ev["g_2"]
        out: List[LinearCombination] = [(l, "beta", [(1, l)]) for l in ("z_b", "g_1", "t")]
        out.append(("outer_sumcheck", "beta", [
            (1, "mask_poly"), (r_alpha_at_beta * (eta[0] + eta[2] * z_b) % p, "z_a"),
            (r_alpha_at_beta * eta[1] * z_b % p, None), (-t * vxb % p, "w"), (-t * x_at_beta % p, None),
            (-vhb % p, "h_1"), (-be * g_1 % p, None)]))
        out.append(("g_2", "gamma", [(1, "g_2")]))
        for m in "abc":
            out.append((m + "_denom", "gamma", [(al * be % p, None), (-al % p, m + "_row"), (-be % p, m + "_col"),
                                                (1, m + "_row_col")]))
        den = [ev[m + "_denom"] for m in "abc"]
        v = vha * vhb % p
        inner = []
        for i, m in enumerate("abc"):
            others = [den[j] for j in range(3) if j != i]
            inner.append((eta[i] * others[0] * others[1] * v % p, m + "_val"))
        b_expr = (ga * g_2 + t * pow(pk.K.n, -1, p)) % p
        inner += [(-(den[0] * den[1] * den[2]) * b_expr % p, None), (-vkg % p, "h_2")]
        out.append(("inner_sumcheck", "gamma", inner))
        return out


class MarlinSNARK:
    """`Marlin::<F, FSF, MarlinKZG10<E, DensePolynomial<F>>, FS, MC>` with MC::FOR_RECURSION = true"""

    def __init__(self, ctx: L.Context, pairing: int):
        self.ctx, self.pairing = ctx, pairing
        self.field = L.SCALAR_FIELD_OF[pairing]

    def index(self, cm: ConstraintMatrices, powers: kzg.Powers, max_degree: int) -> IndexProverKey:
        """Marlin::index: arithmetize and commit to the index polynomials (once per circuit)"""
        if cm.pairing != self.pairing:
            raise ValueError("constraint matrices are over another pairing")
        return IndexProverKey(self.ctx, cm, MarlinKZG10(self.ctx, powers, max_degree))

    def prove(self, pk: IndexProverKey, assignment: np.ndarray, rng: Callable[[int], np.ndarray]) -> Proof:
        """Marlin::prove.  assignment: instance || witness as Montgomery limbs (z[0] = 1)."""
        ctx, field, ops, F = self.ctx, self.field, pk.ops, pk.ops.F
        p = F.p
        H, K, X = pk.H, pk.K, pk.X
        h, k, x = H.n, K.n, X.n
        z = np.ascontiguousarray(assignment, dtype=np.uint64).reshape(-1, 5)
        if z.shape[0] != pk.num_inputs_orig + pk.num_witness:
            raise ValueError("assignment has %d entries, the index expects %d" % (z.shape[0], pk.num_inputs_orig + pk.num_witness))
        polys: Dict[str, LabeledPolynomial] = {l: LabeledPolynomial(l, d) for l, d in pk.index_polys()}
        fs = FiatShamirAlgebraicSpongeRng(field, BASE_FIELD_OF[self.pairing])
        fs.absorb_bytes(PROTOCOL_NAME)
        fs.absorb_native_field_elements([pk.vk_hash])
        public = [F.dec(z[i]) for i in range(pk.num_inputs_orig)] + [0] * (x - pk.num_inputs_orig)
        fs.absorb_nonnative_field_elements(public)

        def commit_round(lps: List[LabeledPolynomial]) -> List[Commitment]:
            comms = pk.pc.commit(lps, rng)
            flat: List[int] = []
            for lp, c in zip(lps, comms):
                polys[lp.label] = lp
                flat += point_to_field_elements(self.pairing, c.comm)
                if c.shifted_comm is not None:
                    flat += point_to_field_elements(self.pairing, c.shifted_comm)
            fs.absorb_native_field_elements(flat)
            return comms

        def add_vanishing_multiple(poly_h: DVec, b: int) -> DVec:
            """poly + b (X^|H| - 1) for a polynomial of |H| coefficients"""
            out = poly_h.resized(h + 1)
            ops.scalar(SUB, out.view(0, 1), out.view(0, 1), b)
            ops.scalar(ADD, out.view(h, h + 1), out.view(h, h + 1), b)
            return out

        # ---- prover_first_round ----
        full_host = np.zeros((h, 5), dtype=np.uint64)
        full_host[:pk.num_inputs_orig] = z[:pk.num_inputs_orig]
        full_host[x:x + pk.num_witness] = z[pk.num_inputs_orig:]
        full = DVec.from_host(ctx, field, full_host)  # formatted input | witness | zero padding, |H| long
        z_a = ops.matvec(pk.csr[0], full, pk.n).resized(h)
        z_b = ops.matvec(pk.csr[1], full, pk.n).resized(h)
        x_poly = ops.ifft(X, full.view(0, x))
        x_evals = ops.fft(H, x_poly)
        w_evals = ops.gather(full, pk.w_index.ptr, h)
        x_masked = ops.gather(x_evals, pk.x_index.ptr, h)
        ops.binary(SUB, w_evals, w_evals, x_masked, h)
        ops.ntt(w_evals, H, True)
        blind = [F.dec(x) for x in draw_many(rng, field, ZK_BOUND)]
        w_full = add_vanishing_multiple(w_evals, blind[0])
        w_poly, w_rem = ops.divide_by_vanishing(w_full, x)
        ops.ntt(z_a, H, True)
        z_a_poly = add_vanishing_multiple(z_a, F.dec(draw_many(rng, field, ZK_BOUND)[0]))
        ops.ntt(z_b, H, True)
        z_b_poly = add_vanishing_multiple(z_b, F.dec(draw_many(rng, field, ZK_BOUND)[0]))
        mask_len = 3 * h + 2 * ZK_BOUND - 2
        mask_host = draw_many(rng, field, mask_len).copy()
        fix = sum(F.dec(mask_host[i]) for i in range(0, mask_len, h)) % p
        mask_host[0] = F.enc(F.dec(mask_host[0]) - fix)  # the mask sums to zero over H
        mask_poly = DVec.from_host(ctx, field, mask_host)
        first = commit_round([LabeledPolynomial("w", w_poly, None, 1), LabeledPolynomial("z_a", z_a_poly, None, 1),
                              LabeledPolynomial("z_b", z_b_poly, None, 1), LabeledPolynomial("mask_poly", mask_poly, None, 1)])
        alpha, eta_a, eta_b, eta_c = fs.squeeze_nonnative_field_elements(4)
        eta = (eta_a, eta_b, eta_c)
        # ---- prover_second_round ----
        vha = H.vanishing_at(alpha)
        if vha == 0:
            raise ArithmeticError("alpha landed in H")
        r_alpha = DVec(ctx, field, h)
        ops.scalar(RSUB, r_alpha, pk.h_elems, alpha)  # alpha - x
        ops.inverse_inplace(r_alpha)
        ops.scalar(MUL, r_alpha, r_alpha, vha)  # u_H(alpha, x) = v_H(alpha) / (alpha - x)
        t_evals = DVec.zeros(ctx, field, h)
        for e, csr_t in zip(eta, pk.csr_t):
            ops.axpy(t_evals, e, ops.matvec(csr_t, r_alpha, h))
        ops.ntt(t_evals, H, True)
        t_poly = t_evals
        z_poly = DVec.zeros(ctx, field, w_poly.n + x)  # w v_X + x_poly
        ops.binary(ADD, z_poly.view(x), z_poly.view(x), w_poly, w_poly.n)
        ops.binary(SUB, z_poly, z_poly, w_poly, w_poly.n)
        ops.binary(ADD, z_poly, z_poly, x_poly, x)
        summed = ops.poly_mul(z_a_poly, z_b_poly)
        ops.scalar(MUL, summed, summed, eta_c)
        ops.axpy(summed, eta_a, z_a_poly)
        ops.axpy(summed, eta_b, z_b_poly)
        ops.ntt(r_alpha, H, True)  # r_alpha_poly
        q_1 = ops.add(mask_poly, ops.poly_mul(r_alpha, summed))
        q_1 = ops.sub(q_1, ops.poly_mul(t_poly, z_poly))
        h_1, x_g_1 = ops.divide_by_vanishing(q_1, h)
        g_1 = x_g_1.view(1, h)
        second = commit_round([LabeledPolynomial("t", t_poly, None, None), LabeledPolynomial("g_1", g_1, h - 2, 1),
                               LabeledPolynomial("h_1", h_1, None, None)])
        (beta,) = fs.squeeze_nonnative_field_elements(1)
        vhb = H.vanishing_at(beta)
        if vhb == 0:
            raise ArithmeticError("beta landed in H")
        # ---- prover_third_round ----
        v = vha * vhb % p
        f_evals = DVec.zeros(ctx, field, k)
        tmp, tmp2 = DVec(ctx, field, k), DVec(ctx, field, k)
        for e, m in zip(eta, pk.arith):
            ops.scalar(RSUB, tmp, m.evals_on_K_row, beta)
            ops.scalar(RSUB, tmp2, m.evals_on_K_col, alpha)
            ops.mul_inplace(tmp, tmp2)
            ops.inverse_inplace(tmp)
            ops.mul_inplace(tmp, m.evals_on_K_val)
            ops.axpy(f_evals, e * v % p, tmp)
        ops.ntt(f_evals, K, True)
        f = f_evals
        g_2 = f.view(1, k)
        # a(X) - b(X) f(X) on a domain that holds its 4|K| - 3 coefficients: 3 denominators, 3 val's and f -> 7 FFTs
        D = Domain(field, 4 * k - 3)
        den_polys, den_ev, val_ev = [], [], []
        for m in pk.arith:
            d = ops.scaled(m.row, -alpha % p)
            ops.axpy(d, -beta % p, m.col)
            ops.binary(ADD, d, d, m.row_col, k)
            ops.scalar(ADD, d.view(0, 1), d.view(0, 1), alpha * beta % p)
            den_polys.append(d)
            den_ev.append(ops.fft(D, d))
            val_ev.append(ops.fft(D, m.val))
        acc = DVec.zeros(ctx, field, D.n)
        for i in range(3):
            j, l = [u for u in range(3) if u != i]
            ops.mul_inplace(val_ev[i], den_ev[j])
            ops.mul_inplace(val_ev[i], den_ev[l])
            ops.axpy(acc, eta[i] * v % p, val_ev[i])
        b_ev = den_ev[0]
        ops.mul_inplace(b_ev, den_ev[1])
        ops.mul_inplace(b_ev, den_ev[2])
        ops.mul_inplace(b_ev, ops.fft(D, f))
        ops.binary(SUB, acc, acc, b_ev, D.n)
        ops.ntt(acc, D, True)
        acc.n = 4 * k - 3
        h_2, rem_k = ops.divide_by_vanishing(acc, k)
        third = commit_round([LabeledPolynomial("g_2", g_2, k - 2, None), LabeledPolynomial("h_2", h_2, None, None)])
        (gamma,) = fs.squeeze_nonnative_field_elements(1)
        ch = dict(alpha=alpha, eta_a=eta_a, eta_b=eta_b, eta_c=eta_c, beta=beta, gamma=gamma)
        # ---- evaluations, linear combinations, batched opening ----
        ev = {"z_b": ops.evaluate(z_b_poly, beta), "t": ops.evaluate(t_poly, beta), "g_1": ops.evaluate(g_1, beta),
              "g_2": ops.evaluate(g_2, gamma)}
        for m, d in zip("abc", den_polys):
            ev[m + "_denom"] = ops.evaluate(d, gamma)
        lcs = AHPForR1CS.construct_linear_combinations(pk, ch, ev, ops.evaluate(x_poly, beta))
        evaluations = sorted((l, ev[l]) for l, _, _ in lcs if l not in LC_WITH_ZERO_EVAL)
        fs.absorb_nonnative_field_elements([e for _, e in evaluations])
        n_open = sum(2 if l in ("g_1", "g_2") else 1 for l, _, _ in lcs)
        opening = fs.squeeze_128_bits_nonnative_field_elements(n_open)
        pc_proof, combined = self._open_combinations(pk, polys, lcs, ch, opening)
        # what a checker needs beyond the proof (tests / bench gate): the polynomials behind the commitments
        self.last_trace = dict(challenges=ch, opening_challenges=opening, polys=polys, lcs=lcs, combined=combined)
        return Proof(self.pairing, [first, second, third], evaluations, pc_proof)

    def _open_combinations(self, pk: IndexProverKey, polys, lcs, ch, opening):
        """MarlinKZG10::open_combinations -> batch_open: per query point the LCs (sorted by label) are folded with
        consecutive opening challenges; a degree-bounded polynomial adds its witness polynomial shifted by
        X^(max_degree - bound) under the next challenge; one kzg10::Proof per point"""
        ops, F, ctx, field = pk.ops, pk.ops.F, self.ctx, self.field
        p, D = F.p, pk.pc.max_degree
        out, combined = [], {}
        for point_label in sorted({pl for _, pl, _ in lcs}):
            zpt = ch[point_label]
            here = sorted((lc for lc in lcs if lc[1] == point_label), key=lambda t: t[0])
            longest = max(polys[l].polynomial.n for _, _, terms in here for _, l in terms if l is not None)
            acc = DVec.zeros(ctx, field, longest)
            acc_r: List[int] = []
            shifted: List[Tuple[int, int, DVec]] = []  # (challenge, shift, witness polynomial)
            shifted_r: List[int] = []
            counter = 0
            for label, _, terms in here:
                cj = opening[counter]
                counter += 1
                named = [(c, l) for c, l in terms if l is not None]
                for c, l in named:
                    ops.axpy(acc, cj * c % p, polys[l].polynomial)
                    if polys[l].rand is not None:
                        acc_r = _host_axpy(F, acc_r, cj * c % p, polys[l].rand)
                lp = polys[named[0][1]]
                if len(named) == 1 and lp.degree_bound is not None:
                    cj1 = opening[counter]
                    counter += 1
                    shifted.append((cj1, D - lp.degree_bound, ops.divide_by_linear(lp.polynomial, zpt)))
                    if lp.shifted_rand is not None:
                        shifted_r = _host_axpy(F, shifted_r, cj1, lp.shifted_rand)
            wit = ops.divide_by_linear(acc, zpt)
            w_len = max([wit.n] + [s + wv.n for _, s, wv in shifted])
            w_poly = wit.resized(w_len)
            for cj1, s, wv in shifted:
                ops.axpy(w_poly.view(s), cj1, wv)
            # the blinding side is hiding_bound + 2 coefficients long: host integers
            rw = _host_div_linear(p, acc_r, zpt)
            rw = _host_add(p, rw, _host_div_linear(p, shifted_r, zpt))
            hiding = bool(acc_r) or bool(shifted_r)
            random_v = (_host_eval(p, acc_r, zpt) + _host_eval(p, shifted_r, zpt)) % p if hiding else None
            w = pk.pc._msm(w_poly, 0, F.enc_many(rw) if rw else None)
            out.append((point_label, w, random_v))
            combined[point_label] = (acc, acc_r, shifted_r, w_poly, rw)
        return out, combined


# blinding polynomials are three coefficients long: they live on the host as integers
def _host_axpy(F: _Fld, acc: List[int], c: int, limbs: np.ndarray) -> List[int]:
    vals = [F.dec(x) for x in np.asarray(limbs).reshape(-1, 5)]
    n = max(len(acc), len(vals))
    return [((acc[i] if i < len(acc) else 0) + c * (vals[i] if i < len(vals) else 0)) % F.p for i in range(n)]


def _host_add(p, a, b):
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % p for i in range(n)]


def _host_eval(p, a, z):
    acc = 0
    for c in reversed(a):
        acc = (acc * z + c) % p
    return acc


def _host_div_linear(p, a, z):
    q = [0] * max(len(a) - 1, 0)
    acc = 0
    for j in range(len(a) - 1, 0, -1):
        acc = (a[j] + z * acc) % p
        q[j - 1] = acc
    return q
