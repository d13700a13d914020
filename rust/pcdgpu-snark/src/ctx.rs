// UNCOMPILED SOURCE (see ../../README.md).
//! One `pcdgpu_ctx` per host thread (the ABI's rule: a context is never shared between threads) and the cache of
//! device-resident keys.  The reference calls `SNARK::prove(&pk, circuit, rng)` with the same `pk` for every node of
//! a PCD (`mod.rs:171,179`), and with a key rebuilt from a fixed-seed rng for the default circuits
//! (`data_structures.rs:135-143`): keys are therefore cached by CONTENT (a hash of their serialized bytes), uploaded
//! -- and their window tables built -- once, and reused by every later proof on this thread.
use crate::GpuSnarkError;
use pcdgpu_sys::*;
use std::cell::RefCell;
use std::collections::HashMap;
use std::ffi::CStr;
use std::os::raw::c_int;

pub struct Ctx {
    pub raw: *mut pcdgpu_ctx,
    /// key fingerprint -> resident (key, matrices); freed with the context
    pub groth16: HashMap<[u8; 32], Resident<pcdgpu_pk>>,
    pub gm17: HashMap<[u8; 32], Resident<pcdgpu_gm17_pk>>,
}
pub struct Resident<K> {
    pub pk: *mut K,
    pub r1cs: *mut pcdgpu_r1cs,
    /// (num_constraints, num_instance, num_witness, nnz a, nnz b, nnz c): a key is reused only with the same shape
    pub shape: (usize, usize, usize, usize, usize, usize),
}

impl Ctx {
    fn new() -> Result<Self, GpuSnarkError> {
        let device: c_int = std::env::var("PCDGPU_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut raw: *mut pcdgpu_ctx = std::ptr::null_mut();
        let rc = unsafe { pcdgpu_ctx_create(device, &mut raw) };
        if rc != PCDGPU_OK {
            // no CPU fallback: the error surfaces to the caller of SNARK::prove
            let message = unsafe { CStr::from_ptr(pcdgpu_strerror(rc)) }.to_string_lossy().into_owned();
            return Err(GpuSnarkError::Backend { code: rc, message });
        }
        Ok(Ctx { raw, groth16: HashMap::new(), gm17: HashMap::new() })
    }
    pub fn check(&self, rc: c_int) -> Result<(), GpuSnarkError> {
        if rc == PCDGPU_OK {
            return Ok(());
        }
        let mut message = unsafe { CStr::from_ptr(pcdgpu_last_error(self.raw)) }.to_string_lossy().into_owned();
        if message.is_empty() {
            message = unsafe { CStr::from_ptr(pcdgpu_strerror(rc)) }.to_string_lossy().into_owned();
        }
        Err(GpuSnarkError::Backend { code: rc, message })
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe {
            for (_, r) in self.groth16.drain() {
                pcdgpu_pk_free(r.pk);
                pcdgpu_r1cs_free(r.r1cs);
            }
            for (_, r) in self.gm17.drain() {
                pcdgpu_gm17_pk_free(r.pk);
                pcdgpu_r1cs_free(r.r1cs);
            }
            pcdgpu_ctx_destroy(self.raw);
        }
    }
}

thread_local! {
    static CTX: RefCell<Option<Ctx>> = RefCell::new(None);
}

/// Runs `f` with this thread's context, creating it on first use.
pub fn with_ctx<T>(f: impl FnOnce(&mut Ctx) -> Result<T, GpuSnarkError>) -> Result<T, GpuSnarkError> {
    CTX.with(|cell| {
        let mut slot = cell.borrow_mut();
        if slot.is_none() {
            *slot = Some(Ctx::new()?);
        }
        f(slot.as_mut().unwrap())
    })
}
