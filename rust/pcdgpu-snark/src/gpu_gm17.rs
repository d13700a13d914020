// UNCOMPILED SOURCE (see ../../README.md).
//! `GpuGM17<E>`: the second SNARK the reference plugs into `ECCyclePCD` (`tests/mnt4_gm17.rs:27-28`,
//! `tests/mnt4_mix_*.rs`).  Same shape as `GpuGroth16`; the R1CS crosses the boundary and the SAP is derived on the GPU.
use crate::ctx::{with_ctx, Resident};
use crate::pack::{self, MontLimbs};
use crate::{GpuSnarkError, PcdGpuPairing};
use ark_crypto_primitives::snark::{CircuitSpecificSetupSNARK, SNARK};
use ark_ff::{Field, UniformRand};
use ark_gm17::{PreparedVerifyingKey, Proof, ProvingKey, VerifyingKey, GM17};
use ark_relations::r1cs::{ConstraintSynthesizer, ConstraintSystem, OptimizationGoal, SynthesisError};
use ark_std::marker::PhantomData;
use ark_std::rand::{CryptoRng, RngCore};
use pcdgpu_sys::*;
use std::os::raw::c_void;

pub struct GpuGM17<E: PcdGpuPairing>(PhantomData<E>);

impl<E> SNARK<E::Fr> for GpuGM17<E>
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
        GM17::<E>::circuit_specific_setup(circuit, rng).map_err(GpuSnarkError::from)
    }

    fn prove<C: ConstraintSynthesizer<E::Fr>, R: RngCore + CryptoRng>(
        pk: &Self::ProvingKey,
        circuit: C,
        rng: &mut R,
    ) -> Result<Self::Proof, Self::Error> {
        // create_random_proof's order: d1, d2, r
        let d1 = E::Fr::rand(rng);
        let d2 = E::Fr::rand(rng);
        let r = E::Fr::rand(rng);
        let cs = ConstraintSystem::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        circuit.generate_constraints(cs.clone())?;
        cs.finalize();
        let matrices = cs.to_matrices().ok_or(SynthesisError::AssignmentMissing)?;
        let z: Vec<u64> = {
            let inner = cs.borrow().ok_or(SynthesisError::MissingCS)?;
            let mut z = Vec::new();
            for x in inner.instance_assignment.iter().chain(inner.witness_assignment.iter()) {
                z.extend_from_slice(x.mont_limbs());
            }
            z
        };
        let (d1l, d2l, rl) = (pack::repr_limbs(&d1), pack::repr_limbs(&d2), pack::repr_limbs(&r));
        let key = pack::fingerprint(pk);
        let shape = pack::shape(&matrices);
        with_ctx(|ctx| {
            if !ctx.gm17.contains_key(&key) {
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
                let ni = matrices.num_instance_variables;
                let nsap = ni + matrices.num_witness_variables + matrices.num_constraints + ni - 1;
                let one = |p: &E::G1Affine| pack::pack_points(core::slice::from_ref(p), 5);
                let one2 = |p: &E::G2Affine| pack::pack_points(core::slice::from_ref(p), g2l);
                let (aq, c1q, c2q, gq) = (pack::pack_points(&pk.a_query, 5), pack::pack_points(&pk.c_query_1, 5),
                                          pack::pack_points(&pk.c_query_2, 5), pack::pack_points(&pk.g_gamma2_z_t, 5));
                let bq = pack::pack_points(&pk.b_query, g2l);
                let (g_gamma_z, g_ab_gamma_z, g_gamma2_z2) = (one(&pk.g_gamma_z), one(&pk.g_ab_gamma_z), one(&pk.g_gamma2_z2));
                let h_gamma_z = one2(&pk.h_gamma_z);
                let mut dpk: *mut pcdgpu_gm17_pk = std::ptr::null_mut();
                let rc = unsafe {
                    pcdgpu_gm17_pk_upload(ctx.raw, E::PAIRING_ID, nsap, ni, pk.g_gamma2_z_t.len(),
                                          aq.as_ptr() as *const c_void, bq.as_ptr() as *const c_void,
                                          c1q.as_ptr() as *const c_void, c2q.as_ptr() as *const c_void,
                                          gq.as_ptr() as *const c_void, g_gamma_z.as_ptr() as *const c_void,
                                          h_gamma_z.as_ptr() as *const c_void, g_ab_gamma_z.as_ptr() as *const c_void,
                                          g_gamma2_z2.as_ptr() as *const c_void, 1, &mut dpk)
                };
                if rc != PCDGPU_OK {
                    unsafe { pcdgpu_r1cs_free(r1cs) };
                    ctx.check(rc)?;
                }
                ctx.gm17.insert(key, Resident { pk: dpk, r1cs, shape });
            }
            let res = &ctx.gm17[&key];
            let mut out = vec![0u64; E::PROOF_AFFINE_BYTES / 8];
            ctx.check(unsafe {
                pcdgpu_gm17_prove(ctx.raw, res.pk, res.r1cs, z.as_ptr() as *const c_void, d1l.as_ptr() as *const c_void,
                                  d2l.as_ptr() as *const c_void, rl.as_ptr() as *const c_void, out.as_mut_ptr() as *mut c_void)
            })?;
            let (a, b, c) = pack::unpack_proof_points::<E>(&out, E::G2_COORD_LIMBS);
            Ok(Proof { a, b, c })
        })
    }

    fn process_vk(vk: &Self::VerifyingKey) -> Result<Self::ProcessedVerifyingKey, Self::Error> {
        GM17::<E>::process_vk(vk).map_err(GpuSnarkError::from)
    }
    fn verify_with_processed_vk(pvk: &Self::ProcessedVerifyingKey, x: &[E::Fr], proof: &Self::Proof) -> Result<bool, Self::Error> {
        GM17::<E>::verify_with_processed_vk(pvk, x, proof).map_err(GpuSnarkError::from)
    }
}
impl<E> CircuitSpecificSetupSNARK<E::Fr> for GpuGM17<E>
where
    E: PcdGpuPairing,
    E::Fr: MontLimbs,
    <E::Fq as Field>::BasePrimeField: MontLimbs,
    <E::Fqe as Field>::BasePrimeField: MontLimbs,
{
}
