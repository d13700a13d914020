// UNCOMPILED SOURCE (see ../../README.md).
//! The gadget catch of SURVEY.md 8(b): `ECCyclePCDConfig` requires `MainSNARKGadget: SNARKGadget<MainField, HelpField,
//! Self::MainSNARK>` (`/root/reference/src/ec_cycle_pcd/mod.rs:31-32`) and upstream implements `SNARKGadget` for
//! `Groth16VerifierGadget<E, P>` against the concrete type `Groth16<E>`.  `GpuGroth16<E>` reuses upstream's key, proof
//! and verifying-key types, so the in-circuit verifier is upstream's, re-exposed under a newtype that names
//! `GpuGroth16<E>` as its SNARK.  Every associated type and function delegates.
use crate::{GpuGM17, GpuGroth16, PcdGpuPairing};
use ark_crypto_primitives::snark::constraints::SNARKGadget;
use ark_ec::PairingEngine;
use ark_gm17::constraints::GM17VerifierGadget;
use ark_groth16::constraints::Groth16VerifierGadget;
use ark_groth16::Groth16;
use ark_r1cs_std::pairing::PairingVar;
use ark_relations::r1cs::SynthesisError;
use ark_std::marker::PhantomData;

pub struct GpuGroth16VerifierGadget<E: PairingEngine, P: PairingVar<E>>(PhantomData<(E, P)>);

impl<E, P> SNARKGadget<E::Fr, E::Fq, GpuGroth16<E>> for GpuGroth16VerifierGadget<E, P>
where
    E: PcdGpuPairing,
    P: PairingVar<E>,
    GpuGroth16<E>: ark_crypto_primitives::snark::SNARK<
        E::Fr,
        ProvingKey = ark_groth16::ProvingKey<E>,
        VerifyingKey = ark_groth16::VerifyingKey<E>,
        Proof = ark_groth16::Proof<E>,
        ProcessedVerifyingKey = ark_groth16::PreparedVerifyingKey<E>,
    >,
{
    type ProcessedVerifyingKeyVar = <Groth16VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, Groth16<E>>>::ProcessedVerifyingKeyVar;
    type VerifyingKeyVar = <Groth16VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, Groth16<E>>>::VerifyingKeyVar;
    type InputVar = <Groth16VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, Groth16<E>>>::InputVar;
    type ProofVar = <Groth16VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, Groth16<E>>>::ProofVar;
    type VerifierSize = <Groth16VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, Groth16<E>>>::VerifierSize;

    fn verifier_size(vk: &ark_groth16::VerifyingKey<E>) -> Self::VerifierSize {
        <Groth16VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, Groth16<E>>>::verifier_size(vk)
    }
    fn verify_with_processed_vk(
        pvk: &Self::ProcessedVerifyingKeyVar,
        x: &Self::InputVar,
        proof: &Self::ProofVar,
    ) -> Result<ark_r1cs_std::boolean::Boolean<E::Fq>, SynthesisError> {
        <Groth16VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, Groth16<E>>>::verify_with_processed_vk(pvk, x, proof)
    }
    fn verify(
        vk: &Self::VerifyingKeyVar,
        x: &Self::InputVar,
        proof: &Self::ProofVar,
    ) -> Result<ark_r1cs_std::boolean::Boolean<E::Fq>, SynthesisError> {
        <Groth16VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, Groth16<E>>>::verify(vk, x, proof)
    }
}

pub struct GpuGM17VerifierGadget<E: PairingEngine, P: PairingVar<E>>(PhantomData<(E, P)>);

impl<E, P> SNARKGadget<E::Fr, E::Fq, GpuGM17<E>> for GpuGM17VerifierGadget<E, P>
where
    E: PcdGpuPairing,
    P: PairingVar<E>,
    GpuGM17<E>: ark_crypto_primitives::snark::SNARK<
        E::Fr,
        ProvingKey = ark_gm17::ProvingKey<E>,
        VerifyingKey = ark_gm17::VerifyingKey<E>,
        Proof = ark_gm17::Proof<E>,
        ProcessedVerifyingKey = ark_gm17::PreparedVerifyingKey<E>,
    >,
{
    type ProcessedVerifyingKeyVar = <GM17VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, ark_gm17::GM17<E>>>::ProcessedVerifyingKeyVar;
    type VerifyingKeyVar = <GM17VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, ark_gm17::GM17<E>>>::VerifyingKeyVar;
    type InputVar = <GM17VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, ark_gm17::GM17<E>>>::InputVar;
    type ProofVar = <GM17VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, ark_gm17::GM17<E>>>::ProofVar;
    type VerifierSize = <GM17VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, ark_gm17::GM17<E>>>::VerifierSize;

    fn verifier_size(vk: &ark_gm17::VerifyingKey<E>) -> Self::VerifierSize {
        <GM17VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, ark_gm17::GM17<E>>>::verifier_size(vk)
    }
    fn verify_with_processed_vk(
        pvk: &Self::ProcessedVerifyingKeyVar,
        x: &Self::InputVar,
        proof: &Self::ProofVar,
    ) -> Result<ark_r1cs_std::boolean::Boolean<E::Fq>, SynthesisError> {
        <GM17VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, ark_gm17::GM17<E>>>::verify_with_processed_vk(pvk, x, proof)
    }
    fn verify(
        vk: &Self::VerifyingKeyVar,
        x: &Self::InputVar,
        proof: &Self::ProofVar,
    ) -> Result<ark_r1cs_std::boolean::Boolean<E::Fq>, SynthesisError> {
        <GM17VerifierGadget<E, P> as SNARKGadget<E::Fr, E::Fq, ark_gm17::GM17<E>>>::verify(vk, x, proof)
    }
}
