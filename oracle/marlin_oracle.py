"""CPU restatement of the Marlin prover the reference binds as MainSNARK / HelpSNARK in its Marlin configuration
(/root/reference/tests/mnt4_marlin.rs:53-94: MarlinSNARK<F, FSF, MarlinKZG10<E, DensePolynomial<F>>,
FiatShamirAlgebraicSpongeRng<F, CF, PoseidonSponge<CF>>, MarlinConfig { FOR_RECURSION = true }>), reached through
IC::MainSNARK::prove / IC::HelpSNARK::prove (/root/reference/src/ec_cycle_pcd/mod.rs:171,179).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU leg may import this file.

PARITY UNPINNED.  ark-marlin and ark-poly-commit are un-vendored, un-pinned git dependencies on their `constraints`
branches (/root/reference/Cargo.toml:41-42); the reference holds no Marlin vectors (its test only asserts
`verify`), and there is no Rust toolchain here.  What follows restates the published algorithms -- ark-marlin
src/ahp/{indexer,prover,verifier,mod}.rs, src/fiat_shamir/{mod,poseidon/mod}.rs, src/lib.rs (`Marlin::prove`);
ark-poly-commit src/marlin_pc/mod.rs and src/kzg10/mod.rs -- from their structure as recalled (SURVEY.md B.7 rates
this "low confidence").  Items that are stand-ins rather than restatements are marked STAND-IN below.  What pins this
file instead is COMPLETENESS AND SOUNDNESS CHECKED END TO END: `check_proof` replays the Fiat-Shamir transcript,
evaluates the AHP verifier's two sumcheck identities from the proof's evaluations, and checks every commitment and
every opening of the batched MarlinKZG10 proof in the exponent for an SRS with a KNOWN trapdoor (beta, gamma).

Polynomials are Python int lists (plain integers mod p, lowest degree first)."""
from dataclasses import dataclass, field as dc_field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import pcd_oracle as o

M32 = 0xFFFFFFFF
M64 = (1 << 64) - 1


# ----------------------------------------------------------------------------------------------------------------
# rand_chacha's ChaChaRng (= ChaCha20, 64-bit block counter, stream 0) seeded with rand_core's seed_from_u64
# (PCG32 expansion), and ark-ff's `Fp320::rand` -- used for the Poseidon round constants only.
# ----------------------------------------------------------------------------------------------------------------
class ChaCha20Rng:
    def __init__(self, seed32: bytes):
        assert len(seed32) == 32
        self.key = [int.from_bytes(seed32[4 * i:4 * i + 4], "little") for i in range(8)]
        self.counter = 0
        self.buf: List[int] = []

    @classmethod
    def seed_from_u64(cls, state: int) -> "ChaCha20Rng":
        seed = b""
        for _ in range(8):
            state = (state * 6364136223846793005 + 11634580027462260723) & M64
            xorshifted = ((((state >> 18) ^ state) >> 27)) & M32
            rot = state >> 59
            x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & M32
            seed += x.to_bytes(4, "little")
        return cls(seed)

    def _block(self):
        c = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574]
        s = c + self.key + [self.counter & M32, (self.counter >> 32) & M32, 0, 0]
        x = list(s)

        def qr(a, b, c_, d):
            x[a] = (x[a] + x[b]) & M32; x[d] ^= x[a]; x[d] = ((x[d] << 16) | (x[d] >> 16)) & M32
            x[c_] = (x[c_] + x[d]) & M32; x[b] ^= x[c_]; x[b] = ((x[b] << 12) | (x[b] >> 20)) & M32
            x[a] = (x[a] + x[b]) & M32; x[d] ^= x[a]; x[d] = ((x[d] << 8) | (x[d] >> 24)) & M32
            x[c_] = (x[c_] + x[d]) & M32; x[b] ^= x[c_]; x[b] = ((x[b] << 7) | (x[b] >> 25)) & M32

        for _ in range(10):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        self.buf = [(x[i] + s[i]) & M32 for i in range(16)]
        self.counter += 1

    def next_u32(self) -> int:
        if not self.buf:
            self._block()
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)

    def field(self, fp: o.PrimeFieldParams) -> int:
        """ark-ff `impl Distribution<Fp320<P>> for Standard`: five u64 limbs, top limb masked to 298 bits, taken AS the
        Montgomery representation, rejected when >= p."""
        while True:
            v = 0
            for i in range(5):
                v |= self.next_u64() << (64 * i)
            v &= (1 << o.MODULUS_BITS) - 1
            if v < fp.p:
                return v * pow(fp.R, -1, fp.p) % fp.p


# ----------------------------------------------------------------------------------------------------------------
# ark-marlin src/fiat_shamir/poseidon/mod.rs: PoseidonSponge<CF>
# ----------------------------------------------------------------------------------------------------------------
POSEIDON_FULL_ROUNDS, POSEIDON_PARTIAL_ROUNDS, POSEIDON_ALPHA = 8, 31, 17
POSEIDON_RATE, POSEIDON_CAPACITY = 2, 1
POSEIDON_MDS = [[1, 0, 1], [1, 1, 0], [0, 1, 1]]
_ARK_CACHE: Dict[str, List[List[int]]] = {}


def poseidon_ark(fp: o.PrimeFieldParams) -> List[List[int]]:
    """PoseidonSponge::new(): round constants F::rand from ChaChaRng::seed_from_u64(123456789), 3 per round."""
    if fp.name not in _ARK_CACHE:
        rng = ChaCha20Rng.seed_from_u64(123456789)
        _ARK_CACHE[fp.name] = [[rng.field(fp) for _ in range(POSEIDON_RATE + POSEIDON_CAPACITY)]
                               for _ in range(POSEIDON_FULL_ROUNDS + POSEIDON_PARTIAL_ROUNDS)]
    return _ARK_CACHE[fp.name]


class PoseidonSponge:
    def __init__(self, fp: o.PrimeFieldParams):
        self.fp, self.p = fp, fp.p
        self.ark = poseidon_ark(fp)
        self.state = [0] * (POSEIDON_RATE + POSEIDON_CAPACITY)
        self.mode = ("absorbing", 0)

    def permute(self):
        p, st = self.p, self.state
        half = POSEIDON_FULL_ROUNDS // 2
        for rnd in range(POSEIDON_FULL_ROUNDS + POSEIDON_PARTIAL_ROUNDS):
            st = [(x + c) % p for x, c in zip(st, self.ark[rnd])]
            if rnd < half or rnd >= half + POSEIDON_PARTIAL_ROUNDS:
                st = [pow(x, POSEIDON_ALPHA, p) for x in st]
            else:
                st[0] = pow(st[0], POSEIDON_ALPHA, p)
            st = [sum(st[j] * POSEIDON_MDS[i][j] for j in range(3)) % p for i in range(3)]
        self.state = st

    def _absorb_internal(self, start: int, elems: Sequence[int]):
        elems = list(elems)
        while True:
            if start + len(elems) <= POSEIDON_RATE:
                for i, e in enumerate(elems):
                    self.state[start + i] = (self.state[start + i] + e) % self.p
                self.mode = ("absorbing", start + len(elems))
                return
            take = POSEIDON_RATE - start
            for i, e in enumerate(elems[:take]):
                self.state[start + i] = (self.state[start + i] + e) % self.p
            self.permute()
            elems = elems[take:]
            start = 0

    def _squeeze_internal(self, start: int, n: int) -> List[int]:
        out: List[int] = []
        while True:
            if start + (n - len(out)) <= POSEIDON_RATE:
                k = n - len(out)
                out += self.state[start:start + k]
                self.mode = ("squeezing", start + k)
                return out
            out += self.state[start:POSEIDON_RATE]
            self.permute()
            start = 0

    def absorb(self, elems: Sequence[int]):
        if not len(elems):
            return
        kind, idx = self.mode
        if kind == "absorbing":
            if idx == POSEIDON_RATE:
                self.permute()
                idx = 0
            self._absorb_internal(idx, elems)
        else:
            self.permute()
            self._absorb_internal(0, elems)

    def squeeze(self, n: int) -> List[int]:
        if n == 0:
            return []
        kind, idx = self.mode
        if kind == "absorbing":
            self.permute()
            return self._squeeze_internal(0, n)
        if idx == POSEIDON_RATE:
            self.permute()
            idx = 0
        return self._squeeze_internal(idx, n)


# ----------------------------------------------------------------------------------------------------------------
# ark-marlin src/fiat_shamir/mod.rs: FiatShamirAlgebraicSpongeRng<F, CF, S>  (F = the SNARK's scalar field, absorbed as
# non-native limbs; CF = the curve's base field = the sponge field; commitments are native)
# ----------------------------------------------------------------------------------------------------------------
#: STAND-IN: ark-nonnative-field's `get_params(298, 298, OptimizationType::Weight)` (a cost search) is not restated;
#: limbs are most significant first as in `get_limbs_representations`
NONNATIVE_NUM_LIMBS, NONNATIVE_BITS_PER_LIMB = 10, 30


def nonnative_limbs(v: int) -> List[int]:
    mask = (1 << NONNATIVE_BITS_PER_LIMB) - 1
    return [(v >> (NONNATIVE_BITS_PER_LIMB * i)) & mask for i in reversed(range(NONNATIVE_NUM_LIMBS))]


class FiatShamirRng:
    def __init__(self, f_fp: o.PrimeFieldParams, cf_fp: o.PrimeFieldParams):
        self.f, self.cf = f_fp, cf_fp
        self.s = PoseidonSponge(cf_fp)

    def absorb_native(self, elems: Sequence[int]):
        self.s.absorb([e % self.cf.p for e in elems])

    def absorb_nonnative(self, elems: Sequence[int]):
        """push_elements_to_sponge + compress_elements: limbs in normal form carry `bits_per_limb + 1` bits each; two
        neighbours are packed into one sponge element when they fit under CF::size_in_bits() - 1."""
        capacity = o.MODULUS_BITS - 1
        per = NONNATIVE_BITS_PER_LIMB + 1
        limbs: List[int] = []
        for e in elems:
            limbs += nonnative_limbs(e % self.f.p)
        dest, i = [], 0
        while i < len(limbs):
            if i + 1 < len(limbs) and 2 * per <= capacity:
                dest.append(limbs[i] * (1 << per) + limbs[i + 1])
                i += 2
            else:
                dest.append(limbs[i])
                i += 1
        self.s.absorb(dest)

    def absorb_bytes(self, data: bytes):
        """bytes -> bits (LSB of each byte first) -> chunks of CF::size_in_bits() - 128 bits -> one element per chunk,
        the first bit of a chunk being the most significant (BigInteger::from_bits is big endian)."""
        cap = o.MODULUS_BITS - 128
        bits = [(b >> i) & 1 for b in data for i in range(8)]
        elems = []
        for k in range(0, len(bits), cap):
            v = 0
            for bit in bits[k:k + cap]:
                v = (v << 1) | bit
            elems.append(v)
        self.s.absorb(elems)

    def _bits(self, num_bits: int) -> List[int]:
        per = o.MODULUS_BITS - 1
        out: List[int] = []
        for e in self.s.squeeze((num_bits + per - 1) // per):
            out += [(e >> i) & 1 for i in reversed(range(per))]  # into_repr().to_bits() big endian, top bits skipped
        return out[:num_bits]

    def _elements(self, n: int, bits_each: int) -> List[int]:
        bits = self._bits(n * bits_each)
        return [sum(b << i for i, b in enumerate(bits[k * bits_each:(k + 1) * bits_each])) % self.f.p for k in range(n)]

    def squeeze_nonnative(self, n: int) -> List[int]:
        return self._elements(n, o.MODULUS_BITS - 1)

    def squeeze_128_bits_nonnative(self, n: int) -> List[int]:
        return self._elements(n, 128)

    def squeeze_native(self, n: int) -> List[int]:
        return self.s.squeeze(n)


# ----------------------------------------------------------------------------------------------------------------
# dense polynomials
# ----------------------------------------------------------------------------------------------------------------
def p_trim(a: List[int]) -> List[int]:
    while a and a[-1] == 0:
        a = a[:-1]
    return a


def p_add(p, a, b, cb: int = 1):
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) + cb * (b[i] if i < len(b) else 0)) % p for i in range(n)]


def p_scale(p, a, c):
    return [x * c % p for x in a]


def p_mul(p, a, b):
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] = (out[i + j] + x * y) % p
    return out


def p_eval(p, a, z):
    acc = 0
    for c in reversed(a):
        acc = (acc * z + c) % p
    return acc


def p_div_vanishing(p, a, n):
    """DensePolynomial::divide_by_vanishing_poly: a = q (X^n - 1) + r"""
    if len(a) <= n:
        return [], list(a)
    q = list(a[n:])
    for i in range(len(q) - n - 1, -1, -1):  # q[i] = a[i+n] + q[i+n]
        q[i] = (q[i] + q[i + n]) % p
    r = [(a[i] + (q[i] if i < len(q) else 0)) % p for i in range(n)]
    return q, r


def p_div_linear(p, a, z):
    q = [0] * max(len(a) - 1, 0)
    acc = 0
    for j in range(len(a) - 1, 0, -1):
        acc = (a[j] + z * acc) % p
        q[j - 1] = acc
    return q, ((a[0] + z * acc) % p if a else 0)


def p_mul_fft(fp, a, b):
    """product through GeneralEvaluationDomain::new(deg + 1), like `&a * &b` on DensePolynomial"""
    if not a or not b:
        return []
    n = len(a) + len(b) - 1
    if n <= 64:
        return p_mul(fp.p, a, b)
    d = o.domain_new(fp, n)
    if d.kind != "radix2":
        return p_mul(fp.p, a, b)
    ea, eb = o.domain_fft(d, a), o.domain_fft(d, b)
    return o.domain_ifft(d, [x * y % fp.p for x, y in zip(ea, eb)])[:n]


# ----------------------------------------------------------------------------------------------------------------
# ark-marlin src/ahp/indexer.rs + constraint_systems.rs
# ----------------------------------------------------------------------------------------------------------------
def reindex_by_subdomain(h: int, x: int, index: int) -> int:
    """EvaluationDomain::reindex_by_subdomain: the first |X| variables sit on the subgroup X of H"""
    period = h // x
    if index < x:
        return index * period
    i = index - x
    return i + i // (period - 1) + 1


def domain_elements(d: o.Domain) -> List[int]:
    out, w = [], 1
    for _ in range(d.size):
        out.append(w)
        w = w * d.omega % d.p
    return out


def vanishing_at(d: o.Domain, x: int) -> int:
    return (pow(x, d.size, d.p) - 1) % d.p


def unnormalized_lagrange_same(d: o.Domain) -> List[int]:
    """batch_eval_unnormalized_bivariate_lagrange_poly_with_same_inputs: u_H(x, x) = |H| x^(|H|-1) for x in H"""
    return [d.size * pow(x, d.size - 1, d.p) % d.p for x in domain_elements(d)]


def unnormalized_lagrange_diff(d: o.Domain, alpha: int) -> List[int]:
    """batch_eval_unnormalized_bivariate_lagrange_poly_with_diff_inputs: u_H(alpha, x) = v_H(alpha) / (alpha - x)"""
    v = vanishing_at(d, alpha)
    return [v * pow((alpha - x) % d.p, -1, d.p) % d.p for x in domain_elements(d)]


@dataclass
class MatrixArithmetization:
    row: List[int]
    col: List[int]
    val: List[int]
    row_col: List[int]
    evals_row: List[int]
    evals_col: List[int]
    evals_val: List[int]


@dataclass
class Index:
    fp: o.PrimeFieldParams
    num_constraints: int  # = num_variables after make_matrices_square
    num_inputs: int  # formatted (padded to a power of two), includes the leading 1
    num_inputs_orig: int
    num_vars_orig: int
    num_non_zero: int
    H: o.Domain
    K: o.Domain
    X: o.Domain
    matrices: List[List[List[Tuple[int, int]]]]  # A, B, C rows of (coeff, column) after padding
    arith: List[MatrixArithmetization]

    def index_polys(self) -> List[Tuple[str, List[int]]]:
        out = []
        for name, m in zip("abc", self.arith):
            out += [(name + "_row", m.row), (name + "_col", m.col), (name + "_val", m.val), (name + "_row_col", m.row_col)]
        return out


def next_pow2(n: int) -> int:
    return 1 << max(n - 1, 0).bit_length()


def index(r1cs: o.R1CS) -> Index:
    """AHPForR1CS::index: pad the instance to a power of two (pad_input_for_indexer_and_prover), make the matrices
    square (make_matrices_square), arithmetize each matrix over K (arithmetize_matrix: "we are dealing with the
    transpose of M" -- row(k) is the VARIABLE's domain element, col(k) the CONSTRAINT's, val(k) = M / u_H(row, row))."""
    fp, p = r1cs.fp, r1cs.fp.p
    ni = next_pow2(r1cs.num_inputs)
    shift = ni - r1cs.num_inputs
    mats = [[[(c, j if j < r1cs.num_inputs else j + shift) for c, j in row] for row in M] for M in (r1cs.A, r1cs.B, r1cs.C)]
    nv = ni + r1cs.num_witness
    nc = r1cs.num_constraints
    n = max(nv, nc)
    for M in mats:
        M += [[] for _ in range(n - nc)]
    nnz = max(sum(len(row) for row in M) for M in mats)
    H, K, X = o.domain_new(fp, n), o.domain_new(fp, nnz), o.domain_new(fp, ni)
    assert H.size % X.size == 0 and H.size > X.size
    helems = domain_elements(H)
    uxx = unnormalized_lagrange_same(H)
    arith = []
    for M in mats:
        rows, cols, vals = [], [], []
        for r, row in enumerate(M):
            for c, j in sorted(row, key=lambda t: t[1]):
                vi = reindex_by_subdomain(H.size, X.size, j)
                rows.append(helems[vi])
                cols.append(helems[r])
                vals.append(c * pow(uxx[vi], -1, p) % p)
        pad = K.size - len(rows)
        rows += [rows[-1]] * pad
        cols += [cols[-1]] * pad
        vals += [0] * pad
        rc = [a * b % p for a, b in zip(rows, cols)]
        arith.append(MatrixArithmetization(o.domain_ifft(K, rows), o.domain_ifft(K, cols), o.domain_ifft(K, vals),
                                           o.domain_ifft(K, rc), rows, cols, vals))
    return Index(fp, n, ni, r1cs.num_inputs, r1cs.num_vars, nnz, H, K, X, mats, arith)


def format_assignment(idx: Index, z: Sequence[int]) -> Tuple[List[int], List[int]]:
    """(formatted public input incl. the leading 1 padded with zeros, witness padded with zeros to the square size)"""
    x = list(z[:idx.num_inputs_orig]) + [0] * (idx.num_inputs - idx.num_inputs_orig)
    w = list(z[idx.num_inputs_orig:])
    w += [0] * (idx.num_constraints - idx.num_inputs - len(w))
    return x, w


# ----------------------------------------------------------------------------------------------------------------
# ark-poly-commit src/marlin_pc/mod.rs over src/kzg10/mod.rs, in the exponent-free form: a commitment is represented by
# the pair of polynomials (p, r) it commits to; `group` turns it into a point through the supplied MSM.
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class Labeled:
    label: str
    poly: List[int]
    degree_bound: Optional[int] = None
    hiding_bound: Optional[int] = None
    rand: List[int] = dc_field(default_factory=list)  # blinding polynomial
    shifted_rand: List[int] = dc_field(default_factory=list)


def kzg_blinding_len(hiding_bound: int) -> int:
    """Randomness::rand: degree hiding_bound + 1 -> hiding_bound + 2 coefficients"""
    return hiding_bound + 2


def pc_commit_randomness(polys: Sequence[Labeled], draw: Callable[[], int]):
    """MarlinKZG10::commit's rng draws, in its order: per polynomial the blinding polynomial, then the one of the
    shifted commitment when there is a degree bound"""
    for lp in polys:
        if lp.hiding_bound is not None:
            lp.rand = [draw() for _ in range(kzg_blinding_len(lp.hiding_bound))]
            if lp.degree_bound is not None:
                lp.shifted_rand = [draw() for _ in range(kzg_blinding_len(lp.hiding_bound))]


# ----------------------------------------------------------------------------------------------------------------
# ark-marlin src/ahp/prover.rs + src/lib.rs (Marlin::prove)
# ----------------------------------------------------------------------------------------------------------------
PROTOCOL_NAME = b"MARLIN-2019"
ZK_BOUND = 1


@dataclass
class Proof:
    commitments: List[List[Tuple[str, object, object]]]  # per round: (label, comm, shifted comm or None)
    evaluations: List[Tuple[str, int]]  # sorted by label
    pc_proof: List[Tuple[str, object, Optional[int]]]  # per query point: (point label, w, random_v)


@dataclass
class ProverTrace:
    """everything check_proof needs beyond the proof: the polynomials behind the commitments"""
    polys: Dict[str, Labeled]
    challenges: Dict[str, int]
    opening_challenges: List[int]
    lcs: List[Tuple[str, str, List[Tuple[int, Optional[str]]]]]
    combined: Dict[str, Tuple[List[int], List[int], List[int], List[int]]]


def vk_hash(f_fp, cf_fp, index_comm_coords: Sequence[int]) -> int:
    """compute_vk_hash: a fresh sponge absorbs the index commitments natively and squeezes one native element"""
    fs = FiatShamirRng(f_fp, cf_fp)
    fs.absorb_native(index_comm_coords)
    return fs.squeeze_native(1)[0]


def linear_combinations(idx: Index, ch: Dict[str, int], ev: Dict[str, int], x_at_beta: int):
    """AHPForR1CS::construct_linear_combinations (with the evaluations folded in as the prover and verifier both do);
    each LC = (label, query point label, [(coefficient, polynomial label or None for the constant)])"""
    p = idx.fp.p
    al, be, ga = ch["alpha"], ch["beta"], ch["gamma"]
    ea, eb, ec = ch["eta_a"], ch["eta_b"], ch["eta_c"]
    vha, vhb, vxb = vanishing_at(idx.H, al), vanishing_at(idx.H, be), vanishing_at(idx.X, be)
    vkg = vanishing_at(idx.K, ga)
    r_ab = (vha - vhb) * pow((al - be) % p, -1, p) % p
    zb, t, g1, g2 = ev["z_b"], ev["t"], ev["g_1"], ev["g_2"]
    lcs = [("z_b", "beta", [(1, "z_b")]), ("g_1", "beta", [(1, "g_1")]), ("t", "beta", [(1, "t")])]
    lcs.append(("outer_sumcheck", "beta", [
        (1, "mask_poly"), (r_ab * (ea + ec * zb) % p, "z_a"), (r_ab * eb % p * zb % p, None), ((-t * vxb) % p, "w"),
        ((-t * x_at_beta) % p, None), ((-vhb) % p, "h_1"), ((-be * g1) % p, None)]))
    lcs.append(("g_2", "gamma", [(1, "g_2")]))
    for m in "abc":
        lcs.append((m + "_denom", "gamma", [(al * be % p, None), ((-al) % p, m + "_row"), ((-be) % p, m + "_col"),
                                            (1, m + "_row_col")]))
    da, db, dc = ev["a_denom"], ev["b_denom"], ev["c_denom"]
    v = vha * vhb % p
    b_at = da * db % p * dc % p
    b_expr = (ga * g2 + t * pow(idx.K.size, -1, p)) % p
    lcs.append(("inner_sumcheck", "gamma", [
        (ea * db % p * dc % p * v % p, "a_val"), (eb * da % p * dc % p * v % p, "b_val"),
        (ec * db % p * da % p * v % p, "c_val"), ((-b_at * b_expr) % p, None), ((-vkg) % p, "h_2")]))
    return lcs


LC_WITH_ZERO_EVAL = ("inner_sumcheck", "outer_sumcheck")


def prove(idx: Index, index_comm_coords: Sequence[int], cf_fp: o.PrimeFieldParams, z: Sequence[int],
          draw: Callable[[], int], max_degree: int,
          group: Callable[[List[int], int, List[int]], object], coords: Callable[[object], List[int]]):
    """Marlin::prove (FOR_RECURSION).  `draw()` = one `F::rand(rng)`; `group(coeffs, shift, blinding)` =
    MSM(powers_of_g[shift..], coeffs) + MSM(powers_of_gamma_g, blinding) as an affine point; `coords(point)` = its
    to_field_elements() (x, y, infinity).  Returns (Proof, ProverTrace)."""
    fp, p = idx.fp, idx.fp.p
    H, K, X = idx.H, idx.K, idx.X
    h, k = H.size, K.size
    x_in, w_in = format_assignment(idx, z)
    full = x_in + w_in
    fs = FiatShamirRng(fp, cf_fp)
    fs.absorb_bytes(PROTOCOL_NAME)
    fs.absorb_native([vk_hash(fp, cf_fp, index_comm_coords)])
    fs.absorb_nonnative(x_in)
    polys: Dict[str, Labeled] = {name: Labeled(name, c) for name, c in idx.index_polys()}

    def commit_round(lps: List[Labeled]):
        pc_commit_randomness(lps, draw)
        out, flat = [], []
        for lp in lps:
            polys[lp.label] = lp
            c = group(lp.poly, 0, lp.rand)
            sc = group(lp.poly, max_degree - lp.degree_bound, lp.shifted_rand) if lp.degree_bound is not None else None
            out.append((lp.label, c, sc))
            flat += coords(c) + (coords(sc) if sc is not None else [])
        fs.absorb_native(flat)
        return out

    # ---- first round (prover_first_round) ----
    za = [sum(c * full[j] for c, j in row) % p for row in idx.matrices[0]]
    zb = [sum(c * full[j] for c, j in row) % p for row in idx.matrices[1]]
    za += [0] * (h - len(za))
    zb += [0] * (h - len(zb))
    x_poly = o.domain_ifft(X, x_in)
    x_evals = o.domain_fft(H, x_poly)
    ratio = h // X.size
    w_ext = w_in + [0] * (h - X.size - len(w_in))
    w_evals = [0 if i % ratio == 0 else (w_ext[i - i // ratio - 1] - x_evals[i]) % p for i in range(h)]
    vh = [p - 1] + [0] * (h - 1) + [1]
    w_poly = p_add(p, o.domain_ifft(H, w_evals), p_mul(p, [draw() for _ in range(ZK_BOUND)], vh))
    w_poly, rem = p_div_vanishing(p, w_poly, X.size)
    assert not p_trim(rem)
    za_poly = p_add(p, o.domain_ifft(H, za), p_mul(p, [draw() for _ in range(ZK_BOUND)], vh))
    zb_poly = p_add(p, o.domain_ifft(H, zb), p_mul(p, [draw() for _ in range(ZK_BOUND)], vh))
    mask_deg = 3 * h + 2 * ZK_BOUND - 3
    mask = [draw() for _ in range(mask_deg + 1)]
    mask[0] = (mask[0] - sum(mask[i] for i in range(0, mask_deg + 1, h))) % p  # sum over H becomes zero
    first = commit_round([Labeled("w", w_poly, None, 1), Labeled("z_a", za_poly, None, 1),
                          Labeled("z_b", zb_poly, None, 1), Labeled("mask_poly", mask, None, 1)])
    alpha, eta_a, eta_b, eta_c = fs.squeeze_nonnative(4)
    assert vanishing_at(H, alpha)
    # ---- second round (prover_second_round) ----
    r_alpha = unnormalized_lagrange_diff(H, alpha)
    t_evals = [0] * h
    for eta, M in zip((eta_a, eta_b, eta_c), idx.matrices):
        for r, row in enumerate(M):
            for c, j in row:
                vi = reindex_by_subdomain(h, X.size, j)
                t_evals[vi] = (t_evals[vi] + eta * c % p * r_alpha[r]) % p
    t_poly = o.domain_ifft(H, t_evals)
    vx = [p - 1] + [0] * (X.size - 1) + [1]
    z_poly = p_add(p, p_mul(p, w_poly, vx), x_poly)
    summed = p_add(p, p_add(p, p_scale(p, za_poly, eta_a), p_scale(p, zb_poly, eta_b)),
                   p_scale(p, p_mul_fft(fp, za_poly, zb_poly), eta_c))
    r_alpha_poly = o.domain_ifft(H, r_alpha)
    q1 = p_add(p, p_add(p, mask, p_mul_fft(fp, r_alpha_poly, summed)), p_mul_fft(fp, t_poly, z_poly), -1)
    h1, xg1 = p_div_vanishing(p, q1, h)
    assert xg1[0] == 0, "outer sumcheck: the remainder has a constant term"
    g1 = p_trim(xg1[1:])
    assert len(g1) <= h - 1
    second = commit_round([Labeled("t", t_poly, None, None), Labeled("g_1", g1, h - 2, 1), Labeled("h_1", h1, None, None)])
    (beta,) = fs.squeeze_nonnative(1)
    assert vanishing_at(H, beta)
    # ---- third round (prover_third_round) ----
    vha, vhb = vanishing_at(H, alpha), vanishing_at(H, beta)
    v = vha * vhb % p
    f_evals = [0] * k
    for eta, m in zip((eta_a, eta_b, eta_c), idx.arith):
        for i in range(k):
            den = (beta - m.evals_row[i]) * (alpha - m.evals_col[i]) % p
            f_evals[i] = (f_evals[i] + eta * v % p * m.evals_val[i] % p * pow(den, -1, p)) % p
    f = o.domain_ifft(K, f_evals)
    g2 = p_trim(f[1:])
    den_polys = [p_add(p, p_add(p, p_add(p, [alpha * beta % p], p_scale(p, m.row, (-alpha) % p)),
                                p_scale(p, m.col, (-beta) % p)), m.row_col) for m in idx.arith]
    a_poly: List[int] = []
    for i, (eta, m) in enumerate(zip((eta_a, eta_b, eta_c), idx.arith)):
        others = [den_polys[j] for j in range(3) if j != i]
        a_poly = p_add(p, a_poly, p_scale(p, p_mul_fft(fp, m.val, p_mul_fft(fp, others[0], others[1])), eta * v % p))
    b_poly = p_mul_fft(fp, p_mul_fft(fp, den_polys[0], den_polys[1]), den_polys[2])
    h2, rem = p_div_vanishing(p, p_add(p, a_poly, p_mul_fft(fp, b_poly, f), -1), k)
    assert not p_trim(rem), "inner sumcheck: a - b f is not a multiple of v_K"
    third = commit_round([Labeled("g_2", g2, k - 2, None), Labeled("h_2", p_trim(h2), None, None)])
    (gamma,) = fs.squeeze_nonnative(1)
    ch = dict(alpha=alpha, eta_a=eta_a, eta_b=eta_b, eta_c=eta_c, beta=beta, gamma=gamma)
    # ---- evaluations and the batched opening (Marlin::prove after the third round) ----
    ev = {"z_b": p_eval(p, zb_poly, beta), "t": p_eval(p, t_poly, beta), "g_1": p_eval(p, g1, beta),
          "g_2": p_eval(p, g2, gamma)}
    for m, dpoly in zip("abc", den_polys):
        ev[m + "_denom"] = p_eval(p, dpoly, gamma)
    lcs = linear_combinations(idx, ch, ev, p_eval(p, x_poly, beta))
    evaluations = sorted((label, ev[label]) for label, _, _ in lcs if label not in LC_WITH_ZERO_EVAL)
    fs.absorb_nonnative([e for _, e in evaluations])
    n_open = sum(2 if label in ("g_1", "g_2") else 1 for label, _, _ in lcs)
    opening = fs.squeeze_128_bits_nonnative(n_open)
    pc_proof, combined = open_combinations(fp, polys, lcs, ch, opening, max_degree, group)
    proof = Proof([first, second, third], evaluations, pc_proof)
    return proof, ProverTrace(polys, ch, opening, lcs, combined)


def lc_polynomial(p, polys: Dict[str, Labeled], terms):
    """the polynomial, blinding polynomial, degree bound and shifted blinding polynomial of a linear combination
    (MarlinKZG10::open_combinations: the constant term is left to the evaluations; an LC keeps a degree bound only when
    it is a single polynomial)"""
    poly: List[int] = []
    rand: List[int] = []
    named = [(c, l) for c, l in terms if l is not None]
    for c, label in named:
        poly = p_add(p, poly, p_scale(p, polys[label].poly, c))
        rand = p_add(p, rand, p_scale(p, polys[label].rand, c))
    bound, srand = None, []
    if len(named) == 1 and polys[named[0][1]].degree_bound is not None:
        assert named[0][0] == 1
        bound, srand = polys[named[0][1]].degree_bound, polys[named[0][1]].shifted_rand
    return poly, rand, bound, srand


def open_combinations(fp, polys, lcs, ch, opening, max_degree, group):
    """MarlinKZG10::open_combinations_individual_opening_challenges -> batch_open...: per query point (sorted by point
    label) the LCs queried there (sorted by label) are folded with consecutive opening challenges; a degree-bounded
    polynomial also contributes its witness polynomial shifted by X^(max_degree - bound) under the next challenge."""
    p = fp.p
    out, combined = [], {}
    for point_label in sorted({pl for _, pl, _ in lcs}):
        z = ch[point_label]
        acc_p: List[int] = []
        acc_r: List[int] = []
        shifted_w: List[int] = []
        shifted_r: List[int] = []
        shifted_rw: List[int] = []
        counter = 0
        # NOTE: one challenge counter per query point would restart at 0 upstream too (each point is opened by its own
        # call to open_individual_opening_challenges)
        for label, pl, terms in sorted((l for l in lcs if l[1] == point_label), key=lambda t: t[0]):
            poly, rand, bound, srand = lc_polynomial(p, polys, terms)
            cj = opening[counter]
            counter += 1
            acc_p = p_add(p, acc_p, p_scale(p, poly, cj))
            acc_r = p_add(p, acc_r, p_scale(p, rand, cj))
            if bound is not None:
                cj1 = opening[counter]
                counter += 1
                wit, _ = p_div_linear(p, poly, z)
                shifted_w = p_add(p, shifted_w, p_scale(p, [0] * (max_degree - bound) + wit, cj1))
                shifted_r = p_add(p, shifted_r, p_scale(p, srand, cj1))
                if srand:
                    shifted_rw = p_add(p, shifted_rw, p_scale(p, p_div_linear(p, srand, z)[0], cj1))
        wit, _ = p_div_linear(p, acc_p, z)
        rwit = p_div_linear(p, acc_r, z)[0] if acc_r else []
        w_poly = p_add(p, wit, shifted_w)
        rw_poly = p_add(p, rwit, shifted_rw)
        hiding = bool(acc_r) or bool(shifted_r)
        random_v = (p_eval(p, acc_r, z) + p_eval(p, shifted_r, z)) % p if hiding else None
        out.append((point_label, group(w_poly, 0, rw_poly), random_v))
        combined[point_label] = (acc_p, acc_r, w_poly, rw_poly)
    return out, combined


# ----------------------------------------------------------------------------------------------------------------
# the check that pins this file (and, through byte equality, the GPU prover): AHP verifier identities + every
# commitment and opening in the exponent of a known-trapdoor SRS
# ----------------------------------------------------------------------------------------------------------------
def check_proof(idx: Index, index_comm_coords: Sequence[int], cf_fp, z_public: Sequence[int], proof: Proof,
                trace: ProverTrace, max_degree: int, srs_beta: int, srs_gamma: int,
                point_of_log: Callable[[int], object], coords: Callable[[object], List[int]]) -> bool:
    """1. replays the transcript from the proof alone and requires the prover's challenges;
    2. AHP verifier (verifier.rs / construct_linear_combinations): both sumcheck LCs evaluate to 0 from the proof's
       evaluations and the openings' values;
    3. KZG10 in the exponent: every commitment == [p(beta) + gamma r(beta)] G (shifted: beta^(D - bound) p(beta)), and
       for each query point  log(C) - v - gamma random_v == log(w) (beta - z)  with C folded from the commitments'
       logs exactly as MarlinKZG10::check_combinations folds the points."""
    fp, p = idx.fp, idx.fp.p
    x_in = list(z_public) + [0] * (idx.num_inputs - len(z_public))
    fs = FiatShamirRng(fp, cf_fp)
    fs.absorb_bytes(PROTOCOL_NAME)
    fs.absorb_native([vk_hash(fp, cf_fp, index_comm_coords)])
    fs.absorb_nonnative(x_in)
    logs: Dict[str, Tuple[int, Optional[int]]] = {}
    ch: Dict[str, int] = {}
    squeezes = [("alpha", "eta_a", "eta_b", "eta_c"), ("beta",), ("gamma",)]
    for rnd, names in zip(proof.commitments, squeezes):
        flat = []
        for label, c, sc in rnd:
            lp = trace.polys[label]
            log = (p_eval(p, lp.poly, srs_beta) + srs_gamma * p_eval(p, lp.rand, srs_beta)) % p
            if coords(c) != coords(point_of_log(log)):
                return False
            slog = None
            if lp.degree_bound is not None:
                slog = (pow(srs_beta, max_degree - lp.degree_bound, p) * p_eval(p, lp.poly, srs_beta)
                        + srs_gamma * p_eval(p, lp.shifted_rand, srs_beta)) % p
                if coords(sc) != coords(point_of_log(slog)):
                    return False
            logs[label] = (log, slog)
            flat += coords(c) + (coords(sc) if sc is not None else [])
        fs.absorb_native(flat)
        for name, val in zip(names, fs.squeeze_nonnative(len(names))):
            ch[name] = val
    if ch != trace.challenges:
        return False
    for name, c in idx.index_polys():
        logs[name] = (p_eval(p, c, srs_beta), None)
    ev = dict(proof.evaluations)
    x_at_beta = p_eval(p, o.domain_ifft(idx.X, x_in), ch["beta"])
    lcs = linear_combinations(idx, ch, ev, x_at_beta)
    fs.absorb_nonnative([e for _, e in sorted(ev.items())])
    n_open = sum(2 if label in ("g_1", "g_2") else 1 for label, _, _ in lcs)
    opening = fs.squeeze_128_bits_nonnative(n_open)
    if opening != trace.opening_challenges:
        return False
    pc = {pl: (w, rv) for pl, w, rv in proof.pc_proof}
    for point_label in sorted({pl for _, pl, _ in lcs}):
        zpt = ch[point_label]
        w, rv = pc[point_label]
        acc_p, acc_r, w_poly, rw_poly = trace.combined[point_label]
        wlog = (p_eval(p, w_poly, srs_beta) + srs_gamma * p_eval(p, rw_poly, srs_beta)) % p
        if coords(w) != coords(point_of_log(wlog)):
            return False
        clog, value, counter = 0, 0, 0
        for label, pl, terms in sorted((l for l in lcs if l[1] == point_label), key=lambda t: t[0]):
            const = sum(c for c, l in terms if l is None) % p
            lc_value = ((ev[label] if label not in LC_WITH_ZERO_EVAL else 0) - const) % p
            lc_log = sum(c * logs[l][0] for c, l in terms if l is not None) % p
            cj = opening[counter]
            counter += 1
            clog = (clog + cj * lc_log) % p
            value = (value + cj * lc_value) % p
            named = [(c, l) for c, l in terms if l is not None]
            if len(named) == 1 and logs[named[0][1]][1] is not None:
                cj1 = opening[counter]
                counter += 1
                shift = pow(srs_beta, max_degree - trace.polys[named[0][1]].degree_bound, p)
                # shifted commitment minus the shifted value: [beta^s (p(beta) - v)] G + gamma [r'(beta)] G
                clog = (clog + cj1 * (logs[named[0][1]][1] - shift * lc_value)) % p
        lhs = (clog - value - srs_gamma * (rv or 0)) % p
        if lhs != wlog * (srs_beta - zpt) % p:
            return False
    return True
