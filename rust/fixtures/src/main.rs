// UNCOMPILED SOURCE (see ../../README.md).
//! The missing fixture set: (input, output) pairs of the REAL arkworks implementation of the path, under
//! `ark_std::test_rng()` (`/root/reference/tests/mnt4_groth16.rs:82`), written as hex limbs in the layout of
//! tests/golden/*.json so that tests/test_arkworks_fixtures.py can replay them against the CPU oracle and the GPU.
//!
//!   cargo run --release -p pcd-fixtures -- ../tests/golden/arkworks
//!
//! Files: constants.json (TWO_ADIC_ROOT_OF_UNITY, LARGE_SUBGROUP_ROOT_OF_UNITY, GENERATOR, group generators),
//! ntt.json (fft / ifft / coset_fft / coset_ifft of random vectors, radix-2 on both fields and mixed-radix on q4),
//! msm.json (VariableBaseMSM::multi_scalar_mul, uniform and witness-like scalars, all four groups),
//! groth16.json / gm17.json (key, matrices, assignment, the rng's draws and the serialized proof of a small circuit).
use ark_ec::{msm::VariableBaseMSM, AffineCurve, PairingEngine, ProjectiveCurve};
use ark_ff::{BigInteger, FftField, FftParameters, Field, PrimeField, UniformRand};
use ark_poly::{EvaluationDomain, GeneralEvaluationDomain};
use ark_relations::r1cs::{ConstraintSynthesizer, ConstraintSystem, ConstraintSystemRef, OptimizationGoal, SynthesisError};
use ark_serialize::CanonicalSerialize;
use ark_std::rand::RngCore;
use pcdgpu_snark::pack::{self, MontLimbs};
use std::fmt::Write as _;
use std::fs;

fn hex(limbs: &[u64]) -> String {
    let mut s = String::with_capacity(16 * limbs.len());
    for w in limbs {
        write!(s, "{:016x}", w).unwrap();
    }
    s
}
fn fp_hex<F: PrimeField + MontLimbs>(x: &F) -> String {
    hex(x.mont_limbs())
}
fn vec_hex<F: PrimeField + MontLimbs>(v: &[F]) -> String {
    v.iter().map(fp_hex).collect::<Vec<_>>().join("")
}
fn repr_hex<F: PrimeField>(x: &F) -> String {
    hex(x.into_repr().as_ref())
}
fn points_hex<G: AffineCurve>(v: &[G], coord_limbs: usize) -> String
where
    <G::BaseField as Field>::BasePrimeField: MontLimbs,
{
    hex(&pack::pack_points(v, coord_limbs))
}

/// fields: both Fr of the cycle
fn dump_constants(dir: &str) {
    fn one<F: PrimeField + FftField + MontLimbs>(name: &str) -> String {
        let mut s = format!("\"{}\": {{", name);
        write!(s, "\"two_adic_root_of_unity\": \"{}\", ", fp_hex(&F::two_adic_root_of_unity())).unwrap();
        write!(s, "\"generator\": \"{}\", ", fp_hex(&F::multiplicative_generator())).unwrap();
        write!(s, "\"two_adicity\": {}, ", <F::FftParams as FftParameters>::TWO_ADICITY).unwrap();
        match F::large_subgroup_root_of_unity() {
            Some(w) => write!(s, "\"large_subgroup_root_of_unity\": \"{}\"", fp_hex(&w)).unwrap(),
            None => write!(s, "\"large_subgroup_root_of_unity\": null").unwrap(),
        }
        s.push('}');
        s
    }
    let body = format!("{{{}, {}, \"g1_mnt4\": \"{}\", \"g2_mnt4\": \"{}\", \"g1_mnt6\": \"{}\", \"g2_mnt6\": \"{}\"}}",
                       one::<ark_mnt4_298::Fr>("r4"), one::<ark_mnt6_298::Fr>("q4"),
                       points_hex(&[ark_mnt4_298::G1Affine::prime_subgroup_generator()], 5),
                       points_hex(&[ark_mnt4_298::G2Affine::prime_subgroup_generator()], 10),
                       points_hex(&[ark_mnt6_298::G1Affine::prime_subgroup_generator()], 5),
                       points_hex(&[ark_mnt6_298::G2Affine::prime_subgroup_generator()], 15));
    fs::write(format!("{}/constants.json", dir), body).unwrap();
}

fn dump_ntt<F: PrimeField + FftField + MontLimbs, R: RngCore>(field: u32, sizes: &[usize], rng: &mut R, cases: &mut Vec<String>) {
    for &n in sizes {
        let dom = GeneralEvaluationDomain::<F>::new(n).expect("domain");
        let x: Vec<F> = (0..dom.size()).map(|_| F::rand(rng)).collect();
        let mut s = format!("{{\"field\": {}, \"requested\": {}, \"size\": {}, \"input\": \"{}\"", field, n, dom.size(), vec_hex(&x));
        let mut v = x.clone();
        dom.fft_in_place(&mut v);
        write!(s, ", \"fft\": \"{}\"", vec_hex(&v)).unwrap();
        v = x.clone();
        dom.ifft_in_place(&mut v);
        write!(s, ", \"ifft\": \"{}\"", vec_hex(&v)).unwrap();
        v = x.clone();
        dom.coset_fft_in_place(&mut v);
        write!(s, ", \"coset_fft\": \"{}\"", vec_hex(&v)).unwrap();
        v = x.clone();
        dom.coset_ifft_in_place(&mut v);
        write!(s, ", \"coset_ifft\": \"{}\"}}", vec_hex(&v)).unwrap();
        cases.push(s);
    }
}

fn dump_msm<G: AffineCurve, R: RngCore>(curve: u32, coord_limbs: usize, sizes: &[usize], rng: &mut R, cases: &mut Vec<String>)
where
    <G::BaseField as Field>::BasePrimeField: MontLimbs,
{
    for &n in sizes {
        for witness_like in &[false, true] {
            let bases: Vec<G> = (0..n).map(|_| G::Projective::rand(rng).into_affine()).collect();
            let scalars: Vec<G::ScalarField> = (0..n)
                .map(|i| {
                    if *witness_like && i % 20 < 6 { G::ScalarField::from(0u64) }
                    else if *witness_like && i % 20 < 11 { G::ScalarField::from(1u64) }
                    else { G::ScalarField::rand(rng) }
                })
                .collect();
            let reprs: Vec<_> = scalars.iter().map(|s| s.into_repr()).collect();
            let out = VariableBaseMSM::multi_scalar_mul(&bases, &reprs).into_affine();
            let sc_hex: String = reprs.iter().map(|b| hex(b.as_ref())).collect();
            cases.push(format!("{{\"curve\": {}, \"n\": {}, \"witness_like\": {}, \"bases\": \"{}\", \"scalars\": \"{}\", \"result\": \"{}\"}}",
                               curve, n, witness_like, points_hex(&bases, coord_limbs), sc_hex, points_hex(&[out], coord_limbs)));
        }
    }
}

/// the circuit of tests/synth.py's family in miniature: w_{i+2} = (z_i + 2 z_{i+1}) * z_{i+1}, two public inputs
struct Chain<F: PrimeField> { a: F, b: F, len: usize }
impl<F: PrimeField> ConstraintSynthesizer<F> for Chain<F> {
    fn generate_constraints(self, cs: ConstraintSystemRef<F>) -> Result<(), SynthesisError> {
        use ark_relations::{lc, r1cs::Variable};
        let mut prev = (cs.new_input_variable(|| Ok(self.a))?, self.a);
        let mut cur = (cs.new_input_variable(|| Ok(self.b))?, self.b);
        for _ in 0..self.len {
            let val = (prev.1 + cur.1.double()) * cur.1;
            let w = cs.new_witness_variable(|| Ok(val))?;
            cs.enforce_constraint(lc!() + prev.0 + (F::from(2u64), cur.0), lc!() + cur.0, lc!() + w)?;
            prev = cur;
            cur = (w, val);
        }
        let _ = Variable::One;
        Ok(())
    }
}

fn dump_groth16<E: pcdgpu_snark::PcdGpuPairing>(pairing: u32, len: usize, cases: &mut Vec<String>)
where
    E::Fr: MontLimbs,
    <E::Fq as Field>::BasePrimeField: MontLimbs,
    <E::Fqe as Field>::BasePrimeField: MontLimbs,
{
    use ark_crypto_primitives::snark::{CircuitSpecificSetupSNARK, SNARK};
    use ark_groth16::Groth16;
    let mut rng = ark_std::test_rng();
    let (a, b) = (E::Fr::rand(&mut rng), E::Fr::rand(&mut rng));
    let (pk, _vk) = Groth16::<E>::circuit_specific_setup(Chain { a, b, len }, &mut rng).unwrap();
    // what prove() will draw, in its order: clone the rng and peek
    let mut peek = rng.clone();
    let (r, s) = (E::Fr::rand(&mut peek), E::Fr::rand(&mut peek));
    let proof = Groth16::<E>::prove(&pk, Chain { a, b, len }, &mut rng).unwrap();
    let cs = ConstraintSystem::<E::Fr>::new_ref();
    cs.set_optimization_goal(OptimizationGoal::Constraints);
    Chain { a, b, len }.generate_constraints(cs.clone()).unwrap();
    cs.finalize();
    let m = cs.to_matrices().unwrap();
    let inner = cs.borrow().unwrap();
    let z: Vec<E::Fr> = inner.instance_assignment.iter().chain(inner.witness_assignment.iter()).cloned().collect();
    let mut bytes = Vec::new();
    proof.serialize(&mut bytes).unwrap();
    let csr_json = |mx: &ark_relations::r1cs::Matrix<E::Fr>| {
        let c = pack::csr(mx);
        format!("{{\"ptr\": {:?}, \"col\": {:?}, \"val\": \"{}\"}}", c.ptr, c.col, hex(&c.val))
    };
    let g2l = E::G2_COORD_LIMBS;
    cases.push(format!(
        "{{\"pairing\": {}, \"m\": {}, \"num_inputs\": {}, \"num_witness\": {}, \"A\": {}, \"B\": {}, \"C\": {}, \"z\": \"{}\", \"r\": \"{}\", \"s\": \"{}\", \
         \"pk\": {{\"alpha_g1\": \"{}\", \"beta_g1\": \"{}\", \"delta_g1\": \"{}\", \"beta_g2\": \"{}\", \"delta_g2\": \"{}\", \"a_query\": \"{}\", \
         \"b_g1_query\": \"{}\", \"b_g2_query\": \"{}\", \"h_query\": \"{}\", \"l_query\": \"{}\"}}, \"proof_affine\": \"{}\", \"proof_bytes\": \"{}\"}}",
        pairing, m.num_constraints, m.num_instance_variables, m.num_witness_variables, csr_json(&m.a), csr_json(&m.b), csr_json(&m.c),
        vec_hex(&z), repr_hex(&r), repr_hex(&s),
        points_hex(&[pk.vk.alpha_g1], 5), points_hex(&[pk.beta_g1], 5), points_hex(&[pk.delta_g1], 5),
        points_hex(&[pk.vk.beta_g2], g2l), points_hex(&[pk.vk.delta_g2], g2l), points_hex(&pk.a_query, 5),
        points_hex(&pk.b_g1_query, 5), points_hex(&pk.b_g2_query, g2l), points_hex(&pk.h_query, 5), points_hex(&pk.l_query, 5),
        format!("{}{}{}", points_hex(&[proof.a], 5), points_hex(&[proof.b], g2l), points_hex(&[proof.c], 5)),
        bytes.iter().map(|b| format!("{:02x}", b)).collect::<String>()));
}

fn main() {
    let dir = std::env::args().nth(1).unwrap_or_else(|| "tests/golden/arkworks".to_string());
    fs::create_dir_all(&dir).unwrap();
    let mut rng = ark_std::test_rng();
    dump_constants(&dir);
    let mut ntt = Vec::new();
    dump_ntt::<ark_mnt4_298::Fr, _>(0, &[1, 2, 8, 512, 1 << 12, 1 << 16], &mut rng, &mut ntt);
    dump_ntt::<ark_mnt6_298::Fr, _>(1, &[1, 2, 8, 512, 1 << 12, 1 << 17, (1 << 17) + 1, 7 << 10], &mut rng, &mut ntt);
    fs::write(format!("{}/ntt.json", dir), format!("[{}]", ntt.join(",\n"))).unwrap();
    let mut msm = Vec::new();
    dump_msm::<ark_mnt4_298::G1Affine, _>(0, 5, &[1, 31, 32, 1000, 1 << 12], &mut rng, &mut msm);
    dump_msm::<ark_mnt4_298::G2Affine, _>(1, 10, &[1, 31, 300], &mut rng, &mut msm);
    dump_msm::<ark_mnt6_298::G1Affine, _>(2, 5, &[1, 31, 32, 1000, 1 << 12], &mut rng, &mut msm);
    dump_msm::<ark_mnt6_298::G2Affine, _>(3, 15, &[1, 31, 300], &mut rng, &mut msm);
    fs::write(format!("{}/msm.json", dir), format!("[{}]", msm.join(",\n"))).unwrap();
    let mut g16 = Vec::new();
    dump_groth16::<ark_mnt4_298::MNT4_298>(0, 30, &mut g16);
    dump_groth16::<ark_mnt6_298::MNT6_298>(1, 30, &mut g16);
    dump_groth16::<ark_mnt4_298::MNT4_298>(0, 1000, &mut g16);
    fs::write(format!("{}/groth16.json", dir), format!("[{}]", g16.join(",\n"))).unwrap();
    // GM17: same shape with ark_gm17::{GM17, create_random_proof}; draws d1, d2, r (see pcdgpu-snark/src/gpu_gm17.rs)
    eprintln!("fixtures written to {}", dir);
}
