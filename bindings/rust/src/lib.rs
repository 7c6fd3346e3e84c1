//! SOURCE ONLY -- never compiled in this repository's image (no cargo/rustc).  Batch entry points for
//! `dusk_poseidon` over the B200 engine: `Hash::digest_batch`, `hades::permute_batch`,
//! `encrypt_batch`, `decrypt_batch`, `merkle4_build`, bound to include/poseidon252_b200.h.
//!
//! `BlsScalar` is `#[repr(transparent)]`-compatible with `[u64; 4]` (its `.0`, read directly at
//! dusk-poseidon src/hash.rs:180), so slices of scalars cross the boundary without conversion.
#![allow(non_camel_case_types)]

use core::ffi::{c_char, c_int, c_void};
use dusk_bls12_381::BlsScalar;
use dusk_jubjub::JubJubAffine;
use dusk_poseidon::{Domain, Error};

#[repr(C)]
pub struct p252_ctx {
    _private: [u8; 0],
}

pub const P252_MEM_HOST: c_int = 0;

extern "C" {
    fn p252_create(device: c_int, out: *mut *mut p252_ctx) -> c_int;
    fn p252_destroy(ctx: *mut p252_ctx);
    fn p252_strerror(status: c_int) -> *const c_char;
    fn p252_permute_batch(ctx: *mut p252_ctx, states: *mut BlsScalar, n: usize, flags: c_int) -> c_int;
    fn p252_hash_batch(ctx: *mut p252_ctx, domain: c_int, input: *const BlsScalar, n: usize, in_len: usize,
                       out: *mut BlsScalar, out_len: usize, flags: c_int) -> c_int;
    fn p252_encrypt_batch(ctx: *mut p252_ctx, msg: *const BlsScalar, n: usize, l: usize,
                          secret_uv: *const BlsScalar, nonce: *const BlsScalar, cipher: *mut BlsScalar,
                          flags: c_int) -> c_int;
    fn p252_decrypt_batch(ctx: *mut p252_ctx, cipher: *const BlsScalar, n: usize, l: usize,
                          secret_uv: *const BlsScalar, nonce: *const BlsScalar, msg: *mut BlsScalar,
                          ok: *mut u8, n_failed: *mut usize, flags: c_int) -> c_int;
    fn p252_merkle4_tree_nodes(n_leaves: usize, n_internal: *mut usize, n_levels: *mut c_int) -> c_int;
    fn p252_merkle4_build(ctx: *mut p252_ctx, leaves: *const BlsScalar, n_leaves: usize,
                          nodes_out: *mut BlsScalar, flags: c_int) -> c_int;
}

/// Engine failures that have no dusk_poseidon::Error counterpart.
#[derive(Debug)]
pub enum BatchError {
    Poseidon(Error),
    Engine(c_int),
}

fn status(rc: c_int) -> Result<(), BatchError> {
    // positive codes are dusk_poseidon::Error in declaration order (src/error.rs:11-32)
    match rc {
        0 => Ok(()),
        1 => Err(BatchError::Poseidon(Error::IOPatternViolation)),
        2 => Err(BatchError::Poseidon(Error::InvalidIOPattern)),
        3 => Err(BatchError::Poseidon(Error::TooFewInputElements)),
        4 => Err(BatchError::Poseidon(Error::EncryptionFailed)),
        5 => Err(BatchError::Poseidon(Error::DecryptionFailed)),
        6 => Err(BatchError::Poseidon(Error::InvalidPoint)),
        e => Err(BatchError::Engine(e)),
    }
}

fn domain_code(d: Domain) -> c_int {
    match d {
        Domain::Merkle4 => 0,
        Domain::Merkle2 => 1,
        Domain::Encryption => 2,
        Domain::Other => 3,
    }
}

/// One CUDA device + stream.  Not `Sync`: calls on one engine serialise.
pub struct Engine(*mut p252_ctx);
unsafe impl Send for Engine {}

impl Engine {
    pub fn new(device: i32) -> Result<Self, BatchError> {
        let mut ctx = core::ptr::null_mut();
        status(unsafe { p252_create(device, &mut ctx) })?;
        Ok(Self(ctx))
    }

    /// `hades::permute_batch`: n independent `ScalarPermutation::permute` (states: n x 5, in place).
    pub fn permute_batch(&self, states: &mut [[BlsScalar; 5]]) -> Result<(), BatchError> {
        status(unsafe { p252_permute_batch(self.0, states.as_mut_ptr().cast(), states.len(), P252_MEM_HOST) })
    }

    /// `Hash::digest_batch`: `inputs.len() / in_len` independent `Hash::digest(domain, chunk)`.
    pub fn digest_batch(&self, domain: Domain, inputs: &[BlsScalar], in_len: usize, output_len: usize)
                        -> Result<Vec<BlsScalar>, BatchError> {
        let n = if in_len == 0 { 0 } else { inputs.len() / in_len };
        let ol = if domain == Domain::Other && output_len > 0 { output_len } else { 1 };
        let mut out = vec![BlsScalar::zero(); n * ol];
        status(unsafe {
            p252_hash_batch(self.0, domain_code(domain), inputs.as_ptr(), n, in_len, out.as_mut_ptr(), ol, P252_MEM_HOST)
        })?;
        Ok(out)
    }

    /// `encrypt_batch`: messages n x L, one shared secret point and nonce per message.
    pub fn encrypt_batch(&self, messages: &[BlsScalar], l: usize, secrets: &[JubJubAffine], nonces: &[BlsScalar])
                         -> Result<Vec<BlsScalar>, BatchError> {
        let n = secrets.len();
        let uv: Vec<BlsScalar> = secrets.iter().flat_map(|p| [p.get_u(), p.get_v()]).collect();
        let mut cipher = vec![BlsScalar::zero(); n * (l + 1)];
        status(unsafe {
            p252_encrypt_batch(self.0, messages.as_ptr(), n, l, uv.as_ptr(), nonces.as_ptr(), cipher.as_mut_ptr(),
                               P252_MEM_HOST)
        })?;
        Ok(cipher)
    }

    /// `decrypt_batch`: per-item `Result` like `dusk_poseidon::decrypt`.
    pub fn decrypt_batch(&self, ciphers: &[BlsScalar], l: usize, secrets: &[JubJubAffine], nonces: &[BlsScalar])
                         -> Result<Vec<Result<Vec<BlsScalar>, Error>>, BatchError> {
        let n = secrets.len();
        let uv: Vec<BlsScalar> = secrets.iter().flat_map(|p| [p.get_u(), p.get_v()]).collect();
        let mut msg = vec![BlsScalar::zero(); n * l];
        let mut ok = vec![0u8; n];
        let mut failed = 0usize;
        status(unsafe {
            p252_decrypt_batch(self.0, ciphers.as_ptr(), n, l, uv.as_ptr(), nonces.as_ptr(), msg.as_mut_ptr(),
                               ok.as_mut_ptr(), &mut failed, P252_MEM_HOST)
        })?;
        Ok((0..n)
            .map(|i| if ok[i] != 0 { Ok(msg[i * l..(i + 1) * l].to_vec()) } else { Err(Error::DecryptionFailed) })
            .collect())
    }

    /// Arity-4 tree of `Domain::Merkle4` digests: internal levels bottom-up, root last.
    pub fn merkle4_build(&self, leaves: &[BlsScalar]) -> Result<Vec<BlsScalar>, BatchError> {
        let mut n_internal = 0usize;
        status(unsafe { p252_merkle4_tree_nodes(leaves.len(), &mut n_internal, core::ptr::null_mut()) })?;
        let mut nodes = vec![BlsScalar::zero(); n_internal];
        status(unsafe { p252_merkle4_build(self.0, leaves.as_ptr(), leaves.len(), nodes.as_mut_ptr(), P252_MEM_HOST) })?;
        Ok(nodes)
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { p252_destroy(self.0) }
    }
}

#[allow(dead_code)]
fn strerror(rc: c_int) -> &'static str {
    unsafe { core::ffi::CStr::from_ptr(p252_strerror(rc)).to_str().unwrap_or("?") }
}

#[allow(dead_code)]
fn _unused(_: *mut c_void) {}
