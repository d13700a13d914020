// UNCOMPILED SOURCE (see ../../README.md).
//! Explicit packing between arkworks' in-memory types and the ABI's byte layouts.  `GroupAffine` and `Fp320` are
//! `repr(Rust)`: nothing here transmutes a struct; every field element is written limb by limb
//! (`Fp320(BigInteger320([u64; 5]))` holds the Montgomery representative a*R mod p, R = 2^320 -- exactly the ABI's
//! field element), every point as x || y with the point at infinity as all zeros.
use ark_ec::{AffineCurve, PairingEngine};
use ark_ff::{BigInteger, Field, PrimeField, Zero};
use ark_relations::r1cs::{ConstraintMatrices, Matrix};

/// the five Montgomery limbs of a prime-field element (ark-ff 0.2/0.3: the tuple field `.0` of `Fp320<P>` is the
/// Montgomery representative as a `BigInteger320`)
pub fn push_fp<F: PrimeField>(out: &mut Vec<u64>, x: &F) {
    // `F::BigInt: AsRef<[u64]>`; the representative is reached through the struct's public field in these
    // revisions: `x.0.as_ref()`.  Written through a small trait so the one revision-specific line is isolated.
    out.extend_from_slice(MontLimbs::mont_limbs(x));
}
pub trait MontLimbs {
    fn mont_limbs(&self) -> &[u64];
}
impl<P: ark_ff::Fp320Parameters> MontLimbs for ark_ff::Fp320<P> {
    fn mont_limbs(&self) -> &[u64] {
        (self.0).as_ref()
    }
}
/// the plain-integer limbs of a scalar (`into_repr()`): what the ABI takes for r, s, d1, d2 and MSM scalars
pub fn repr_limbs<F: PrimeField>(x: &F) -> Vec<u64> {
    x.into_repr().as_ref().to_vec()
}
/// base-field coefficients of an extension element, lowest first (Fp2: c0, c1; Fp3: c0, c1, c2)
pub fn push_ext<F: Field>(out: &mut Vec<u64>, x: &F)
where
    F::BasePrimeField: MontLimbs,
{
    for c in x.to_base_prime_field_elements() {
        out.extend_from_slice(c.mont_limbs());
    }
}
/// affine point -> x || y (all zeros for the point at infinity)
pub fn push_affine<G: AffineCurve>(out: &mut Vec<u64>, p: &G, coord_limbs: usize)
where
    <G::BaseField as Field>::BasePrimeField: MontLimbs,
{
    if p.is_zero() {
        out.extend(std::iter::repeat(0u64).take(2 * coord_limbs));
        return;
    }
    let (x, y) = xy(p);
    push_ext(out, &x);
    push_ext(out, &y);
}
/// coordinates of a non-zero short-Weierstrass affine point.  The accessor differs between arkworks revisions
/// (public `.x` / `.y` fields of `GroupAffine` in 0.2 / 0.3, `xy()` later); going through the uncompressed canonical
/// serialization (x || y, both revisions) keeps this crate independent of it.  Pin a revision and replace it by the
/// direct field access if the two extra copies per point ever show up in a profile (keys are packed once per key).
pub fn xy<G: AffineCurve>(p: &G) -> (G::BaseField, G::BaseField) {
    let mut bytes = Vec::new();
    ark_serialize::CanonicalSerialize::serialize_uncompressed(p, &mut bytes).expect("serialize point");
    let half = bytes.len() / 2;
    let x = <G::BaseField as ark_serialize::CanonicalDeserialize>::deserialize(&bytes[..half]).expect("x");
    let y = <G::BaseField as ark_serialize::CanonicalDeserialize>::deserialize(&bytes[half..]).expect("y");
    (x, y)
}
pub fn pack_points<G: AffineCurve>(pts: &[G], coord_limbs: usize) -> Vec<u64>
where
    <G::BaseField as Field>::BasePrimeField: MontLimbs,
{
    let mut out = Vec::with_capacity(pts.len() * 2 * coord_limbs);
    for p in pts {
        push_affine(&mut out, p, coord_limbs);
    }
    out
}

/// `ConstraintMatrices` rows of `(coeff, col)` -> CSR arrays (u32 row_ptr, u32 col, Montgomery values).  Columns are
/// already "instance first, then witness" in ark-relations' matrices.
pub struct Csr {
    pub ptr: Vec<u32>,
    pub col: Vec<u32>,
    pub val: Vec<u64>,
}
pub fn csr<F: PrimeField + MontLimbs>(m: &Matrix<F>) -> Csr {
    let nnz: usize = m.iter().map(|r| r.len()).sum();
    let mut out = Csr { ptr: Vec::with_capacity(m.len() + 1), col: Vec::with_capacity(nnz), val: Vec::with_capacity(5 * nnz) };
    out.ptr.push(0);
    for row in m {
        for (coeff, col) in row {
            out.col.push(*col as u32);
            out.val.extend_from_slice(coeff.mont_limbs());
        }
        out.ptr.push(out.col.len() as u32);
    }
    out
}
pub fn shape<F: PrimeField>(m: &ConstraintMatrices<F>) -> (usize, usize, usize, usize, usize, usize) {
    (m.num_constraints, m.num_instance_variables, m.num_witness_variables, m.a_num_non_zero, m.b_num_non_zero, m.c_num_non_zero)
}

/// A || B || C (x || y each, Montgomery limbs) -> the three affine points of a proof
pub fn unpack_point<G: AffineCurve>(limbs: &[u64]) -> G {
    if limbs.iter().all(|w| *w == 0) {
        return G::zero();
    }
    // the ABI returns Montgomery limbs (a R mod p); rebuild the field elements from them (fp_from_mont)
    let half = limbs.len() / 2;
    let x = ext_from_mont::<G::BaseField>(&limbs[..half]);
    let y = ext_from_mont::<G::BaseField>(&limbs[half..]);
    from_xy::<G>(x, y)
}
fn ext_from_mont<F: Field>(limbs: &[u64]) -> F {
    let k = limbs.len() / 5;
    let mut coeffs = Vec::with_capacity(k);
    for c in 0..k {
        coeffs.push(fp_from_mont::<F::BasePrimeField>(&limbs[5 * c..5 * c + 5]));
    }
    F::from_base_prime_field_elems(&coeffs).expect("coefficient count")
}
/// Montgomery limbs -> field element: from_repr(limbs) gives limbs * R (it multiplies by R^2 and reduces), so undo one R
fn fp_from_mont<F: PrimeField>(limbs: &[u64]) -> F {
    let mut big = F::BigInt::default();
    big.as_mut().copy_from_slice(limbs);
    let as_if_plain = F::from_repr(big).expect("canonical limbs");  // = (limbs) as a plain integer -> Montgomery(limbs)
    // limbs = a*R  =>  from_repr(limbs) represents the VALUE a*R; divide by R = 2^320 mod p
    as_if_plain * F::from(2u64).pow([320u64]).inverse().unwrap()
}
fn from_xy<G: AffineCurve>(x: G::BaseField, y: G::BaseField) -> G {
    let mut bytes = Vec::new();
    ark_serialize::CanonicalSerialize::serialize_uncompressed(&x, &mut bytes).unwrap();
    ark_serialize::CanonicalSerialize::serialize_uncompressed(&y, &mut bytes).unwrap();
    // uncompressed affine = x || y with the infinity flag in y's spare bits (zero here)
    <G as ark_serialize::CanonicalDeserialize>::deserialize_uncompressed(&bytes[..]).expect("point on curve")
}
pub fn unpack_proof_points<E: PairingEngine>(out: &[u64], g2_coord_limbs: usize) -> (E::G1Affine, E::G2Affine, E::G1Affine) {
    let a = unpack_point::<E::G1Affine>(&out[0..10]);
    let b = unpack_point::<E::G2Affine>(&out[10..10 + 2 * g2_coord_limbs]);
    let c = unpack_point::<E::G1Affine>(&out[10 + 2 * g2_coord_limbs..20 + 2 * g2_coord_limbs]);
    (a, b, c)
}

/// content fingerprint of a key (cache key of the resident-key cache): SHA-256-free, dependency-free FNV over the
/// canonical serialization, widened to 32 bytes
pub fn fingerprint<T: ark_serialize::CanonicalSerialize>(t: &T) -> [u8; 32] {
    let mut bytes = Vec::new();
    t.serialize_uncompressed(&mut bytes).expect("serialize key");
    let mut out = [0u8; 32];
    for lane in 0..4u64 {
        let mut h: u64 = 0xcbf29ce484222325 ^ (lane.wrapping_mul(0x9e3779b97f4a7c15));
        for b in &bytes {
            h ^= *b as u64;
            h = h.wrapping_mul(0x100000001b3);
        }
        out[8 * lane as usize..8 * lane as usize + 8].copy_from_slice(&h.to_le_bytes());
    }
    out
}
