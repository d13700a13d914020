// UNCOMPILED SOURCE (see ../../README.md).
//! `GpuGroth16<E>` / `GpuGM17<E>`: drop-in `SNARK` implementations for arkworks-rs/pcd's `MainSNARK` / `HelpSNARK`
//! (`/root/reference/src/ec_cycle_pcd/mod.rs:24-33`).  Keys, proofs and verifying keys are upstream's own types, so
//! setup and verification delegate to the CPU implementation and only `prove` crosses into libpcdgpu.so -- the four
//! `SNARK::prove` calls of a PCD step (`mod.rs:171,179`; `data_structures.rs:139-143,343-350`).
pub mod ctx;
pub mod gadgets;
pub mod gpu_gm17;
pub mod gpu_groth16;
pub mod pack;

pub use gadgets::{GpuGM17VerifierGadget, GpuGroth16VerifierGadget};
pub use gpu_gm17::GpuGM17;
pub use gpu_groth16::GpuGroth16;

use ark_ec::PairingEngine;
use std::os::raw::c_int;

/// The two pairings of the cycle, as libpcdgpu numbers them.
pub trait PcdGpuPairing: PairingEngine {
    const PAIRING_ID: c_int;
    /// bytes of `Proof { a, b, c }` as three affine points x || y: 320 (MNT4-298) / 400 (MNT6-298)
    const PROOF_AFFINE_BYTES: usize;
    /// u64 limbs per G2 coordinate: 10 (Fq2) / 15 (Fq3)
    const G2_COORD_LIMBS: usize;
}
impl PcdGpuPairing for ark_mnt4_298::MNT4_298 {
    const PAIRING_ID: c_int = pcdgpu_sys::PCDGPU_MNT4_298;
    const PROOF_AFFINE_BYTES: usize = 320;
    const G2_COORD_LIMBS: usize = 10;
}
impl PcdGpuPairing for ark_mnt6_298::MNT6_298 {
    const PAIRING_ID: c_int = pcdgpu_sys::PCDGPU_MNT6_298;
    const PROOF_AFFINE_BYTES: usize = 400;
    const G2_COORD_LIMBS: usize = 15;
}

/// Error of the GPU provers: synthesis errors pass through; backend errors carry libpcdgpu's code and message.
#[derive(Debug)]
pub enum GpuSnarkError {
    Synthesis(ark_relations::r1cs::SynthesisError),
    Backend { code: c_int, message: String },
}
impl core::fmt::Display for GpuSnarkError {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        match self {
            GpuSnarkError::Synthesis(e) => write!(f, "{}", e),
            GpuSnarkError::Backend { code, message } => write!(f, "libpcdgpu error {}: {}", code, message),
        }
    }
}
impl std::error::Error for GpuSnarkError {}
impl From<ark_relations::r1cs::SynthesisError> for GpuSnarkError {
    fn from(e: ark_relations::r1cs::SynthesisError) -> Self {
        GpuSnarkError::Synthesis(e)
    }
}
