// UNCOMPILED SOURCE (see ../../README.md).
//! `GpuGroth16<E>`: `Groth16<E>` whose `prove` runs on the GPU.  Replaces `IC::MainSNARK::prove` /
//! `IC::HelpSNARK::prove` of `/root/reference/src/ec_cycle_pcd/mod.rs:171,179` (and the default-circuit proves of
//! `data_structures.rs:139-143,343-350`) through `pcdgpu_groth16_prove`.
use crate::ctx::{with_ctx, Resident};
use crate::pack::{self, MontLimbs};
use crate::{GpuSnarkError, PcdGpuPairing};
use ark_crypto_primitives::snark::{CircuitSpecificSetupSNARK, SNARK};
use ark_ff::{Field, UniformRand};
use ark_groth16::{Groth16, PreparedVerifyingKey, Proof, ProvingKey, VerifyingKey};
use ark_relations::r1cs::{ConstraintSynthesizer, ConstraintSystem, OptimizationGoal, SynthesisError};
use ark_std::marker::PhantomData;
use ark_std::rand::{CryptoRng, RngCore};
use pcdgpu_sys::*;
use std::os::raw::c_void;

pub struct GpuGroth16<E: PcdGpuPairing>(PhantomData<E>);

impl<E> SNARK<E::Fr> for GpuGroth16<E>
where
    E: PcdGpuPairing,
    E::Fr: MontLimbs,
    <E::Fq as Field>::BasePrimeField: MontLimbs,
    <E::Fqe as Field>::BasePrimeField: MontLimbs,
{
    type ProvingKey = ProvingKey<E>;
    type VerifyingKey = VerifyingKey<E>;
    type Proof = Proof<E>;
    type ProcessedVerifyingKey = PreparedVerifyingKey<E>;
    type Error = GpuSnarkError;

    fn circuit_specific_setup<C: ConstraintSynthesizer<E::Fr>, R: RngCore + CryptoRng>(
        circuit: C,
        rng: &mut R,
    ) -> Result<(Self::ProvingKey, Self::VerifyingKey), Self::Error> {
        // one-off, CPU; the key it returns is what prove() uploads
        Groth16::<E>::circuit_specific_setup(circuit, rng).map_err(GpuSnarkError::from)
    }

    fn prove<C: ConstraintSynthesizer<E::Fr>, R: RngCore + CryptoRng>(
        pk: &Self::ProvingKey,
        circuit: C,
        rng: &mut R,
    ) -> Result<Self::Proof, Self::Error> {
        // the draw order of ark-groth16's create_random_proof: r first, then s -- BEFORE synthesis, as upstream does
        let r = E::Fr::rand(rng);
        let s = E::Fr::rand(rng);
        // synthesis stays on the CPU, exactly as in create_proof_with_reduction
        let cs = ConstraintSystem::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        circuit.generate_constraints(cs.clone())?;
        cs.finalize();
        let matrices = cs.to_matrices().ok_or(SynthesisError::AssignmentMissing)?;
        let z: Vec<u64> = {
            let inner = cs.borrow().ok_or(SynthesisError::MissingCS)?;
            let mut z = Vec::with_capacity(5 * (inner.instance_assignment.len() + inner.witness_assignment.len()));
            for x in inner.instance_assignment.iter().chain(inner.witness_assignment.iter()) {
                z.extend_from_slice(x.mont_limbs());
            }
            z
        };
        let (r_limbs, s_limbs) = (pack::repr_limbs(&r), pack::repr_limbs(&s));
        let key = pack::fingerprint(pk);
        let shape = pack::shape(&matrices);

        with_ctx(|ctx| {
            let stale = ctx.groth16.get(&key).map(|res| res.shape != shape).unwrap_or(false);
            if stale {
                let old = ctx.groth16.remove(&key).unwrap();
                unsafe {
                    pcdgpu_pk_free(old.pk);
                    pcdgpu_r1cs_free(old.r1cs);
                }
            }
            if !ctx.groth16.contains_key(&key) {
                // first proof under this key on this thread: upload matrices and key, build the window tables
                let (a, b, c) = (pack::csr(&matrices.a), pack::csr(&matrices.b), pack::csr(&matrices.c));
                let mut r1cs: *mut pcdgpu_r1cs = std::ptr::null_mut();
                ctx.check(unsafe {
                    pcdgpu_r1cs_upload(ctx.raw, E::PAIRING_ID, matrices.num_constraints, matrices.num_instance_variables,
                                       matrices.num_witness_variables,
                                       a.ptr.as_ptr(), a.col.as_ptr(), a.val.as_ptr() as *const c_void,
                                       b.ptr.as_ptr(), b.col.as_ptr(), b.val.as_ptr() as *const c_void,
                                       c.ptr.as_ptr(), c.col.as_ptr(), c.val.as_ptr() as *const c_void, &mut r1cs)
                })?;
                let g2l = E::G2_COORD_LIMBS;
                let one = |p: &E::G1Affine| pack::pack_points(core::slice::from_ref(p), 5);
                let one2 = |p: &E::G2Affine| pack::pack_points(core::slice::from_ref(p), g2l);
                let (alpha, beta1, delta1) = (one(&pk.vk.alpha_g1), one(&pk.beta_g1), one(&pk.delta_g1));
                let (beta2, delta2) = (one2(&pk.vk.beta_g2), one2(&pk.vk.delta_g2));
                let (aq, b1q, hq, lq) = (pack::pack_points(&pk.a_query, 5), pack::pack_points(&pk.b_g1_query, 5),
                                         pack::pack_points(&pk.h_query, 5), pack::pack_points(&pk.l_query, 5));
                let b2q = pack::pack_points(&pk.b_g2_query, g2l);
                let num_vars = matrices.num_instance_variables + matrices.num_witness_variables;
                if pk.a_query.len() != num_vars || pk.b_g1_query.len() != num_vars || pk.b_g2_query.len() != num_vars
                    || pk.l_query.len() != matrices.num_witness_variables {
                    unsafe { pcdgpu_r1cs_free(r1cs) };
                    return Err(GpuSnarkError::Synthesis(SynthesisError::MalformedVerifyingKey));
                }
                let mut dpk: *mut pcdgpu_pk = std::ptr::null_mut();
                let rc = unsafe {
                    pcdgpu_pk_upload(ctx.raw, E::PAIRING_ID, num_vars, matrices.num_instance_variables, pk.h_query.len(),
                                     alpha.as_ptr() as *const c_void, beta1.as_ptr() as *const c_void,
                                     delta1.as_ptr() as *const c_void, beta2.as_ptr() as *const c_void,
                                     delta2.as_ptr() as *const c_void, aq.as_ptr() as *const c_void,
                                     b1q.as_ptr() as *const c_void, b2q.as_ptr() as *const c_void,
                                     hq.as_ptr() as *const c_void, lq.as_ptr() as *const c_void, 1, &mut dpk)
                };
                if rc != PCDGPU_OK {
                    unsafe { pcdgpu_r1cs_free(r1cs) };
                    ctx.check(rc)?;
                }
                ctx.groth16.insert(key, Resident { pk: dpk, r1cs, shape });
            }
            let res = &ctx.groth16[&key];
            let mut out = vec![0u64; E::PROOF_AFFINE_BYTES / 8];
            ctx.check(unsafe {
                pcdgpu_groth16_prove(ctx.raw, res.pk, res.r1cs, z.as_ptr() as *const c_void,
                                     r_limbs.as_ptr() as *const c_void, s_limbs.as_ptr() as *const c_void,
                                     out.as_mut_ptr() as *mut c_void)
            })?;
            let (a, b, c) = pack::unpack_proof_points::<E>(&out, E::G2_COORD_LIMBS);
            Ok(Proof { a, b, c })
        })
    }

    fn process_vk(vk: &Self::VerifyingKey) -> Result<Self::ProcessedVerifyingKey, Self::Error> {
        Groth16::<E>::process_vk(vk).map_err(GpuSnarkError::from)
    }
    fn verify_with_processed_vk(pvk: &Self::ProcessedVerifyingKey, x: &[E::Fr], proof: &Self::Proof) -> Result<bool, Self::Error> {
        Groth16::<E>::verify_with_processed_vk(pvk, x, proof).map_err(GpuSnarkError::from)  // one pairing check, CPU
    }
}
impl<E> CircuitSpecificSetupSNARK<E::Fr> for GpuGroth16<E>
where
    E: PcdGpuPairing,
    E::Fr: MontLimbs,
    <E::Fq as Field>::BasePrimeField: MontLimbs,
    <E::Fqe as Field>::BasePrimeField: MontLimbs,
{
}
