//! src/ffi.rs — the C ABI of libstark_b200 (include/stark_b200.h) as seen from Rust.
//!
//! The `extern "C"` block is generated from the header (tools/gen_rust_ffi.py) and checked against it by
//! tests/test_shim.py; everything below it is the small amount of glue the patched modules share.
#![allow(non_camel_case_types, dead_code)]

use std::collections::HashMap;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_uint, c_void};
use std::sync::{Mutex, OnceLock};

use crate::fields::FieldElement;

macro_rules! opaque {
    ($($name:ident),*) => { $( #[repr(C)] pub struct $name { _private: [u8; 0] } )* };
}
opaque!(stark_ctx, stark_vec, stark_tree, stark_fri, stark_channel, stark_mg, stark_mg_fri);

extern "C" {
    // BEGIN GENERATED (tools/gen_rust_ffi.py)
    pub fn stark_last_error() -> *const c_char;
    pub fn stark_version() -> *const c_char;
    pub fn stark_ctx_create(modulus: u64, generator: u64, device: c_int, out: *mut *mut stark_ctx) -> c_int;
    pub fn stark_ctx_destroy(ctx: *mut stark_ctx);
    pub fn stark_ctx_sync(ctx: *mut stark_ctx) -> c_int;
    pub fn stark_ctx_modulus(ctx: *const stark_ctx) -> u64;
    pub fn stark_ctx_generator(ctx: *const stark_ctx) -> u64;
    pub fn stark_ctx_root_of_unity(ctx: *const stark_ctx, log_n: c_uint) -> u64;
    pub fn stark_ctx_two_adicity(ctx: *const stark_ctx) -> c_uint;
    pub fn stark_ctx_launch_count(ctx: *const stark_ctx) -> u64;
    pub fn stark_ctx_stream(ctx: *const stark_ctx) -> *mut c_void;
    pub fn stark_ctx_set_timing(ctx: *mut stark_ctx, on: c_int) -> c_int;
    pub fn stark_ctx_read_timing(ctx: *mut stark_ctx, ms: *mut f64, units: *mut f64, launches: *mut u64) -> c_int;
    pub fn stark_measure_int_peak(ctx: *mut stark_ctx, tops: *mut f64) -> c_int;
    pub fn stark_measure_pipe_mix(ctx: *mut stark_ctx, steps_per_s: *mut f64) -> c_int;
    pub fn stark_vec_upload(ctx: *mut stark_ctx, host: *const u64, n: usize, out: *mut *mut stark_vec) -> c_int;
    pub fn stark_vec_alloc(ctx: *mut stark_ctx, n: usize, out: *mut *mut stark_vec) -> c_int;
    pub fn stark_vec_download(v: *const stark_vec, offset: usize, n: usize, host: *mut u64) -> c_int;
    pub fn stark_vec_from_device(ctx: *mut stark_ctx, device_u32: *const c_void, n: usize, out: *mut *mut stark_vec) -> c_int;
    pub fn stark_vec_len(v: *const stark_vec) -> usize;
    pub fn stark_vec_device_ptr(v: *const stark_vec) -> *mut c_void;
    pub fn stark_vec_destroy(v: *mut stark_vec);
    pub fn stark_ntt(ctx: *mut stark_ctx, inout: *mut u64, log_n: c_uint) -> c_int;
    pub fn stark_intt(ctx: *mut stark_ctx, inout: *mut u64, log_n: c_uint) -> c_int;
    pub fn stark_coset_evaluate(ctx: *mut stark_ctx, coeffs: *const u64, n_coeffs: usize, log_n: c_uint, offset: u64, out: *mut u64) -> c_int;
    pub fn stark_coset_interpolate(ctx: *mut stark_ctx, evals: *const u64, log_n: c_uint, offset: u64, coeffs_out: *mut u64) -> c_int;
    pub fn stark_coset_lde(ctx: *mut stark_ctx, evals: *const u64, log_n: c_uint, offset_in: u64, log_blowup: c_uint, offset_out: u64, out: *mut u64) -> c_int;
    pub fn stark_batch_inverse(ctx: *mut stark_ctx, inout: *mut u64, n: usize) -> c_int;
    pub fn stark_quotient_pointwise(ctx: *mut stark_ctx, num: *const u64, den: *const u64, n: usize, out: *mut u64) -> c_int;
    pub fn stark_coset_domain(ctx: *mut stark_ctx, log_n: c_uint, offset: u64, out: *mut u64) -> c_int;
    pub fn stark_coset_evaluate_dev(ctx: *mut stark_ctx, coeffs: *const stark_vec, log_n: c_uint, offset: u64, out: *mut *mut stark_vec) -> c_int;
    pub fn stark_coset_interpolate_dev(ctx: *mut stark_ctx, evals: *const stark_vec, offset: u64, out: *mut *mut stark_vec) -> c_int;
    pub fn stark_coset_lde_dev(ctx: *mut stark_ctx, evals: *const stark_vec, offset_in: u64, log_blowup: c_uint, offset_out: u64, out: *mut *mut stark_vec) -> c_int;
    pub fn stark_batch_inverse_dev(ctx: *mut stark_ctx, a: *const stark_vec, out: *mut *mut stark_vec) -> c_int;
    pub fn stark_quotient_pointwise_dev(ctx: *mut stark_ctx, num: *const stark_vec, den: *const stark_vec, out: *mut *mut stark_vec) -> c_int;
    pub fn stark_ntt_batch_dev(ctx: *mut stark_ctx, v: *mut stark_vec, log_m: c_uint, inverse: c_int) -> c_int;
    pub fn stark_pow_mul_dev(ctx: *mut stark_ctx, v: *mut stark_vec, inner_len: usize, outer0: usize, product: c_int, inner_stride: usize, base: u64, c0: u64, log_table: c_uint) -> c_int;
    pub fn stark_peer_alloc(ctx: *mut stark_ctx, n: usize, out: *mut *mut stark_vec, handle: *mut u8) -> c_int;
    pub fn stark_peer_open(ctx: *mut stark_ctx, handle: *const u8, dptr: *mut *mut c_void) -> c_int;
    pub fn stark_peer_close(ctx: *mut stark_ctx, dptr: *mut c_void) -> c_int;
    pub fn stark_fourstep_phase_a(ctx: *mut stark_ctx, coeffs: *const stark_vec, log_n: c_uint, offset: u64, world: c_uint, rank: c_uint, peer_rows: *const *mut c_void, peer_flags: *const *mut c_void, epoch: u32) -> c_int;
    pub fn stark_fourstep_phase_c(ctx: *mut stark_ctx, rows: *mut stark_vec, log_n: c_uint, world: c_uint, rank: c_uint, peer_blocks: *const *mut c_void, peer_flags: *const *mut c_void, epoch: u32) -> c_int;
    pub fn stark_fourstep_wait(ctx: *mut stark_ctx, own_flags: *const c_void, slot: c_uint, world: c_uint, epoch: u32) -> c_int;
    pub fn stark_merkle_commit(ctx: *mut stark_ctx, leaves: *const u64, n: usize, out: *mut *mut stark_tree) -> c_int;
    pub fn stark_merkle_commit_dev(ctx: *mut stark_ctx, leaves: *const stark_vec, out: *mut *mut stark_tree) -> c_int;
    pub fn stark_merkle_root(t: *const stark_tree, root: *mut u8) -> c_int;
    pub fn stark_merkle_root_hex(t: *const stark_tree, out: *mut c_char) -> c_int;
    pub fn stark_merkle_num_leaves(t: *const stark_tree) -> usize;
    pub fn stark_merkle_depth(t: *const stark_tree) -> usize;
    pub fn stark_merkle_open(t: *const stark_tree, idx: usize, path: *mut u8, cap: usize, path_len: *mut usize) -> c_int;
    pub fn stark_merkle_node(t: *const stark_tree, level: usize, j: usize, out: *mut u8) -> c_int;
    pub fn stark_tree_destroy(t: *mut stark_tree);
    pub fn stark_channel_new(modulus: u64, out: *mut *mut stark_channel) -> c_int;
    pub fn stark_channel_destroy(ch: *mut stark_channel);
    pub fn stark_channel_send(ch: *mut stark_channel, msg: *const u8, len: usize) -> c_int;
    pub fn stark_channel_receive_random_field_element(ch: *mut stark_channel, out: *mut u64) -> c_int;
    pub fn stark_channel_receive_random_int(ch: *mut stark_channel, min: u64, max: u64, show_in_proof: c_int, out: *mut u64) -> c_int;
    pub fn stark_channel_proof_size(ch: *const stark_channel) -> usize;
    pub fn stark_channel_compressed_proof_size(ch: *const stark_channel) -> usize;
    pub fn stark_channel_state(ch: *const stark_channel) -> *const c_char;
    pub fn stark_channel_proof_len(ch: *const stark_channel) -> usize;
    pub fn stark_channel_proof_msg(ch: *const stark_channel, i: usize, data: *mut *const u8) -> usize;
    pub fn stark_channel_compressed_len(ch: *const stark_channel) -> usize;
    pub fn stark_channel_compressed_msg(ch: *const stark_channel, i: usize, data: *mut *const u8) -> usize;
    pub fn stark_channel_proof_flat(ch: *const stark_channel, out: *mut u8) -> usize;
    pub fn stark_fri_begin(ctx: *mut stark_ctx, coeffs: *const u64, n_coeffs: usize, log_n: c_uint, offset: u64, out: *mut *mut stark_fri, root: *mut u8) -> c_int;
    pub fn stark_fri_begin_dev(ctx: *mut stark_ctx, coeffs: *const stark_vec, log_n: c_uint, offset: u64, out: *mut *mut stark_fri, root: *mut u8) -> c_int;
    pub fn stark_fri_degree(f: *const stark_fri, degree: *mut i64) -> c_int;
    pub fn stark_fri_fold(f: *mut stark_fri, beta: u64, root: *mut u8) -> c_int;
    pub fn stark_fri_final(f: *const stark_fri, value: *mut u64, final_poly_len: *mut usize) -> c_int;
    pub fn stark_fri_num_layers(f: *const stark_fri) -> usize;
    pub fn stark_fri_layer_len(f: *const stark_fri, k: usize) -> usize;
    pub fn stark_fri_layer_read(f: *const stark_fri, k: usize, offset: usize, n: usize, out: *mut u64) -> c_int;
    pub fn stark_fri_layer_tree(f: *const stark_fri, k: usize) -> *const stark_tree;
    pub fn stark_fri_open(f: *const stark_fri, indices: *const u64, n_idx: usize, out: *mut u8, cap: usize, len: *mut usize) -> c_int;
    pub fn stark_fri_begin_external(ctx: *mut stark_ctx, coeffs: *const stark_vec, log_n: c_uint, offset: u64, layer0: *const stark_vec, root0: *const u8, out: *mut *mut stark_fri) -> c_int;
    pub fn stark_fri_open_layers(f: *const stark_fri, first_layer: usize, indices: *const u64, n_idx: usize, out: *mut u8, cap: usize, len: *mut usize) -> c_int;
    pub fn stark_fri_destroy(f: *mut stark_fri);
    pub fn stark_fri_begin_to_host(ctx: *mut stark_ctx, coeffs: *const u64, n_coeffs: usize, log_n: c_uint, offset: u64, layers_out: *mut u64, cap: usize, out: *mut *mut stark_fri, root: *mut u8) -> c_int;
    pub fn stark_fri_layers_wait(f: *const stark_fri) -> c_int;
    pub fn stark_fri_layer_host_offset(f: *const stark_fri, k: usize) -> usize;
    pub fn stark_fri_commit_to_host(ctx: *mut stark_ctx, coeffs: *const u64, n_coeffs: usize, log_n: c_uint, offset: u64, ch: *mut stark_channel, layers_out: *mut u64, cap: usize, out: *mut *mut stark_fri) -> c_int;
    pub fn stark_fri_commit_to_host_async(ctx: *mut stark_ctx, coeffs: *const u64, n_coeffs: usize, log_n: c_uint, offset: u64, ch: *mut stark_channel, layers_out: *mut u64, cap: usize, out: *mut *mut stark_fri) -> c_int;
    pub fn stark_fri_commit(ctx: *mut stark_ctx, coeffs: *const u64, n_coeffs: usize, log_n: c_uint, offset: u64, ch: *mut stark_channel, out: *mut *mut stark_fri) -> c_int;
    pub fn stark_fri_commit_dev(ctx: *mut stark_ctx, coeffs: *const stark_vec, log_n: c_uint, offset: u64, ch: *mut stark_channel, out: *mut *mut stark_fri) -> c_int;
    pub fn stark_decommit_fri_layers(f: *const stark_fri, index: usize, ch: *mut stark_channel) -> c_int;
    pub fn stark_decommit_fri(f: *const stark_fri, num_queries: usize, max_index: usize, ch: *mut stark_channel) -> c_int;
    pub fn stark_mg_unique_id(id: *mut u8) -> c_int;
    pub fn stark_mg_create(ctx: *mut stark_ctx, id: *const u8, rank: c_uint, world: c_uint, out: *mut *mut stark_mg) -> c_int;
    pub fn stark_mg_adopt(ctx: *mut stark_ctx, nccl_comm: *mut c_void, rank: c_uint, world: c_uint, out: *mut *mut stark_mg) -> c_int;
    pub fn stark_mg_destroy(mg: *mut stark_mg);
    pub fn stark_mg_rank(mg: *const stark_mg) -> c_uint;
    pub fn stark_mg_world(mg: *const stark_mg) -> c_uint;
    pub fn stark_mg_barrier(mg: *mut stark_mg) -> c_int;
    pub fn stark_mg_commit_columns(mg: *mut stark_mg, n_cols: usize, columns: *const *const u64, log_rows: c_uint, offset_in: u64, log_blowup: c_uint, offset_out: u64, roots: *mut u8, ldes: *mut *mut stark_vec, trees: *mut *mut stark_tree) -> c_int;
    pub fn stark_mg_fourstep_lde(mg: *mut stark_mg, coeffs: *const stark_vec, log_n: c_uint, offset: u64, transport: c_int, block: *mut *mut stark_vec) -> c_int;
    pub fn stark_mg_commit_leaf_ranges(mg: *mut stark_mg, block: *const stark_vec, subtree: *mut *mut stark_tree, root: *mut u8, subtree_roots: *mut u8) -> c_int;
    pub fn stark_mg_fri_commit(mg: *mut stark_mg, coeffs: *const stark_vec, log_n: c_uint, offset: u64, transport: c_int, ch: *mut stark_channel, out: *mut *mut stark_mg_fri) -> c_int;
    pub fn stark_mg_decommit_fri(f: *mut stark_mg_fri, num_queries: usize, max_index: usize, ch: *mut stark_channel) -> c_int;
    pub fn stark_mg_stark101_prove(mg: *mut stark_mg, a1: u64, log_trace: c_uint, log_blowup: c_uint, num_queries: usize, transport: c_int, ch: *mut stark_channel) -> c_int;
    pub fn stark_mg_fri_proof(f: *const stark_mg_fri) -> *const stark_fri;
    pub fn stark_mg_fri_subtree(f: *const stark_mg_fri) -> *const stark_tree;
    pub fn stark_mg_fri_destroy(f: *mut stark_mg_fri);
    pub fn stark_merkle_verify(root: *const u8, n_leaves: usize, idx: usize, value: u64, path: *const u8, path_len: usize, ok: *mut c_int) -> c_int;
    pub fn stark_fri_verify(proof_flat: *const u8, proof_len: usize, modulus: u64, generator: u64, log_n: c_uint, offset: u64, num_queries: usize, max_index: usize, log_degree_bound: c_uint, ok: *mut c_int, reason: *mut c_char) -> c_int;
    pub fn stark101_verify(proof_flat: *const u8, proof_len: usize, modulus: u64, generator: u64, claimed_last: u64, log_trace: c_uint, log_blowup: c_uint, num_queries: usize, ok: *mut c_int, reason: *mut c_char) -> c_int;
    pub fn stark101_prove(ctx: *mut stark_ctx, a1: u64, log_trace: c_uint, log_blowup: c_uint, num_queries: usize, ch: *mut stark_channel) -> c_int;
    pub fn stark101_trace_poly(ctx: *mut stark_ctx, a1: u64, log_trace: c_uint, coeffs: *mut *mut stark_vec, last_value: *mut u64) -> c_int;
    pub fn stark101_composition_range(ctx: *mut stark_ctx, f_block: *const stark_vec, start: usize, count: usize, alpha: *const u64, last_value: u64, log_trace: c_uint, log_blowup: c_uint, cp_block: *mut *mut stark_vec) -> c_int;
    // END GENERATED
}

/// The reference's error convention is `panic!` (ops.rs:143, merkle/mod.rs:25); nothing unwinds across the FFI itself.
pub fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(stark_last_error()) }.to_string_lossy().into_owned();
        panic!("libstark_b200 [{}]: {}", rc, msg);
    }
}

struct CtxPtr(*mut stark_ctx);
// a stark_ctx serialises its callers internally (one recursive mutex per context), so the pointer may be shared
unsafe impl Send for CtxPtr {}
unsafe impl Sync for CtxPtr {}

/// One context per (device, MODULUS) for the life of the process.  The device is `STARK_B200_DEVICE` (default 0); the
/// generator of F_p^* is chosen by the library (the smallest one: 5 for p = 3221225473), which fixes
/// w_n = generator^((p-1)/n) for every domain.
pub fn ctx<const M: u64>() -> *mut stark_ctx {
    static CONTEXTS: OnceLock<Mutex<HashMap<u64, CtxPtr>>> = OnceLock::new();
    let map = CONTEXTS.get_or_init(|| Mutex::new(HashMap::new()));
    let mut guard = map.lock().expect("context table poisoned");
    if let Some(c) = guard.get(&M) {
        return c.0;
    }
    let device: c_int = std::env::var("STARK_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
    let mut out: *mut stark_ctx = std::ptr::null_mut();
    check(unsafe { stark_ctx_create(M, 0, device, &mut out) });
    guard.insert(M, CtxPtr(out));
    out
}

/// `&[FieldElement<M>]` as the `*const u64` the ABI takes (FieldElement is #[repr(transparent)] over u64).
pub fn as_u64_ptr<const M: u64>(v: &[FieldElement<M>]) -> *const u64 {
    v.as_ptr() as *const u64
}
pub fn as_u64_mut_ptr<const M: u64>(v: &mut [FieldElement<M>]) -> *mut u64 {
    v.as_mut_ptr() as *mut u64
}

/// If `domain` is a power-of-two coset offset * <w_n> in natural order with w_n the library's root of unity, returns
/// (log2 n, offset).  Checks D[1]/D[0] against the context's root and then every element (n multiplications on the
/// host: cheap next to what the caller is about to avoid).
pub fn as_coset<const M: u64>(domain: &[FieldElement<M>]) -> Option<(c_uint, u64)> {
    let n = domain.len();
    if n == 0 || !n.is_power_of_two() {
        return None;
    }
    let log_n = n.trailing_zeros() as c_uint;
    let c = ctx::<M>();
    if log_n > unsafe { stark_ctx_two_adicity(c) } {
        return None;
    }
    let offset = domain[0];
    if offset == FieldElement::<M>::zero() {
        return None;
    }
    let w = FieldElement::<M>::new(unsafe { stark_ctx_root_of_unity(c, log_n) });
    let mut x = offset;
    for d in domain.iter() {
        if *d != x {
            return None;
        }
        x = x * w;
    }
    Some((log_n, offset.value()))
}
