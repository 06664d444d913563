//! src/fri/fri_commit.rs — the FRI commit phase and its query openings (reference src/fri/fri_commit.rs:72-179) with
//! the per-layer work on the device: layer 0 is one coset NTT + tree, every later layer ONE fused kernel chain that
//! folds the previous layer in evaluation space and hashes the result (`next_fri_layer` + `MerkleTree::new`, :53-65, :97).
//!
//! Repairs to the draft: the types carry `<const M: u64>`; the root enters the channel as the ASCII
//! bytes of its hex string (what fri_verify.rs:24-25 parses), where the draft calls `.to_vec()` on a `String`;
//! `MerkleTree::new` takes the layer by value (merkle/mod.rs:10); `get_authentication_path` exists (merkle/mod.rs here).
//! The loop condition, the order of channel operations and every byte sent are the draft's.
use crate::channel::Channel;
use crate::ffi;
use crate::fields::FieldElement;
use crate::fri::coset_fri::CosetFri;
use crate::merkle::MerkleTree;
use crate::polynomial::Polynomial;

/// Each layer's evaluations, each layer's Merkle tree, the final polynomial (reference :9-13).  `fri_layers` holds the
/// evaluations BY VALUE when the proof was made by `fri_commit` (a copy back from HBM: 8 bytes per element, 2N elements
/// in total) and is empty when it was made by `fri_commit_resident`; `fri_merkles[k]` borrows the k-th tree of `handle`.
pub struct FRIProof<const M: u64> {
    pub fri_layers: Vec<Vec<FieldElement<M>>>,
    pub fri_merkles: Vec<MerkleTree<M>>,
    pub final_poly: Polynomial<M>,
    handle: *mut ffi::stark_fri,
}

unsafe impl<const M: u64> Send for FRIProof<M> {}

impl<const M: u64> FRIProof<M> {
    pub fn num_layers(&self) -> usize {
        unsafe { ffi::stark_fri_num_layers(self.handle) }
    }
    pub fn layer_len(&self, k: usize) -> usize {
        unsafe { ffi::stark_fri_layer_len(self.handle, k) }
    }
    /// `fri_layers[k][offset .. offset + n]` read from the device on demand.
    pub fn read_layer(&self, k: usize, offset: usize, n: usize) -> Vec<FieldElement<M>> {
        let mut out = vec![FieldElement::<M>::zero(); n];
        ffi::check(unsafe { ffi::stark_fri_layer_read(self.handle, k, offset, n, ffi::as_u64_mut_ptr(&mut out)) });
        out
    }
    /// Element offset of layer k in the buffer given to `fri_commit_into` (`None`: the proof was not streamed to the host).
    pub fn layer_host_offset(&self, k: usize) -> Option<usize> {
        let o = unsafe { ffi::stark_fri_layer_host_offset(self.handle, k) };
        if o == usize::MAX { None } else { Some(o) }
    }
    pub(crate) fn raw(&self) -> *const ffi::stark_fri {
        self.handle
    }
}

impl<const M: u64> Drop for FRIProof<M> {
    fn drop(&mut self) {
        self.fri_merkles.clear(); // borrowed trees first
        if !self.handle.is_null() {
            unsafe { ffi::stark_fri_destroy(self.handle) }
        }
    }
}

fn hex_lower(bytes: &[u8; 32]) -> String {
    const DIGITS: &[u8; 16] = b"0123456789abcdef";
    let mut s = String::with_capacity(64);
    for b in bytes.iter() {
        s.push(DIGITS[(b >> 4) as usize] as char);
        s.push(DIGITS[(b & 15) as usize] as char);
    }
    s
}

fn commit_loop<const M: u64>(poly: Polynomial<M>, domain: &CosetFri<M>, channel: &mut Channel<M>,
                             sink: Option<&mut [FieldElement<M>]>) -> (*mut ffi::stark_fri, Polynomial<M>) {
    assert!(domain.domain_size.is_power_of_two(), "FRI domain size must be a power of two");
    let log_n = domain.domain_size.trailing_zeros();
    let c = ffi::ctx::<M>();
    assert_eq!(domain.omega.value(), unsafe { ffi::stark_ctx_root_of_unity(c, log_n) },
               "the domain generator must be the library's root of unity (CosetFri::with_library_root)");
    let mut handle: *mut ffi::stark_fri = std::ptr::null_mut();
    let mut root = [0u8; 32];
    // :78-79  evaluate on the coset, build the tree (with a sink: every layer also streams to it on the copy stream)
    let streaming = sink.is_some();
    match sink {
        None => ffi::check(unsafe {
            ffi::stark_fri_begin(c, ffi::as_u64_ptr(&poly.coefficients), poly.coefficients.len(), log_n, domain.offset.value(), &mut handle,
                                 root.as_mut_ptr())
        }),
        Some(buf) => ffi::check(unsafe {
            ffi::stark_fri_begin_to_host(c, ffi::as_u64_ptr(&poly.coefficients), poly.coefficients.len(), log_n, domain.offset.value(),
                                         ffi::as_u64_mut_ptr(buf), buf.len(), &mut handle, root.as_mut_ptr())
        }),
    }
    channel.send(hex_lower(&root).as_bytes()); // :86
    let mut degree: i64 = 0;
    loop {
        ffi::check(unsafe { ffi::stark_fri_degree(handle, &mut degree) });
        if degree < 1 {
            break; // :89  while poly.degree >= 1
        }
        let beta = channel.receive_random_field_element(); // :91
        ffi::check(unsafe { ffi::stark_fri_fold(handle, beta.value(), root.as_mut_ptr()) }); // :94-97
        channel.send(hex_lower(&root).as_bytes()); // :100
    }
    let (mut value, mut len) = (0u64, 0usize);
    ffi::check(unsafe { ffi::stark_fri_final(handle, &mut value, &mut len) });
    let final_value = FieldElement::<M>::new(value); // zero when the polynomial is zero (:109-113)
    channel.send(&final_value.to_bytes()); // :114
    let final_poly = if len == 0 { Polynomial::zero() } else { Polynomial::new(vec![final_value]) };
    if streaming {
        ffi::check(unsafe { ffi::stark_fri_layers_wait(handle) }); // the caller's buffer is complete when we return
    }
    (handle, final_poly)
}

fn borrow_trees<const M: u64>(handle: *mut ffi::stark_fri) -> Vec<MerkleTree<M>> {
    let layers = unsafe { ffi::stark_fri_num_layers(handle) };
    (0..layers).map(|k| MerkleTree::borrowed(unsafe { ffi::stark_fri_layer_tree(handle, k) })).collect()
}

/// fri_commit (reference :72-122), layers returned by value like the reference does.
pub fn fri_commit<const M: u64>(poly: Polynomial<M>, domain: &CosetFri<M>, channel: &mut Channel<M>) -> FRIProof<M> {
    let (handle, final_poly) = commit_loop(poly, domain, channel, None);
    let layers = unsafe { ffi::stark_fri_num_layers(handle) };
    let mut fri_layers = Vec::with_capacity(layers);
    for k in 0..layers {
        let n = unsafe { ffi::stark_fri_layer_len(handle, k) };
        let mut evals = vec![FieldElement::<M>::zero(); n];
        ffi::check(unsafe { ffi::stark_fri_layer_read(handle, k, 0, n, ffi::as_u64_mut_ptr(&mut evals)) });
        fri_layers.push(evals);
    }
    FRIProof { fri_layers, fri_merkles: borrow_trees(handle), final_poly, handle }
}

/// The same commit phase with the layers left in HBM (`fri_layers` empty; `read_layer` copies ranges on demand): what a
/// prover that only opens a few dozen positions wants -- copying 2N elements back costs more than the whole commit.
pub fn fri_commit_resident<const M: u64>(poly: Polynomial<M>, domain: &CosetFri<M>, channel: &mut Channel<M>) -> FRIProof<M> {
    let (handle, final_poly) = commit_loop(poly, domain, channel, None);
    FRIProof { fri_layers: Vec::new(), fri_merkles: borrow_trees(handle), final_poly, handle }
}

/// The commit phase with every layer ALSO written by value into `layers` (layer k at `proof.layer_host_offset(k)`, u64 per
/// element; `2 * domain_size` elements hold every layer), copied by the library's second stream while the following layers
/// are hashed: the by-value `FRIProof` of the reference at (almost) the cost of `fri_commit_resident`, provided `layers` is
/// page-locked (`cudaHostRegister` / `cudaHostAlloc`); `fri_layers` stays empty, the slice is the caller's.
pub fn fri_commit_into<const M: u64>(poly: Polynomial<M>, domain: &CosetFri<M>, channel: &mut Channel<M>,
                                     layers: &mut [FieldElement<M>]) -> FRIProof<M> {
    let (handle, final_poly) = commit_loop(poly, domain, channel, Some(layers));
    FRIProof { fri_layers: Vec::new(), fri_merkles: borrow_trees(handle), final_poly, handle }
}

/// Decommit all FRI layers for a single query index (reference :137-165), signature as in the reference: elements from
/// the by-value layers, paths from the trees (one device opening per path).
pub fn decommit_fri_layers<const M: u64>(index: usize, fri_layers: &[Vec<FieldElement<M>>], fri_merkles: &[MerkleTree<M>],
                                         channel: &mut Channel<M>) {
    for (layer_evals, merkle_tree) in fri_layers.iter().zip(fri_merkles) {
        let length = layer_evals.len();
        if length == 1 {
            channel.send(&layer_evals[0].to_bytes()); // :147-149 (and, as written, it falls through)
        }
        let idx = index % length; // :152
        let sibling_idx = (idx + length / 2) % length; // :153
        channel.send(&layer_evals[idx].to_bytes()); // :156
        channel.send(&merkle_tree.get_authentication_path(idx)); // :157-158
        channel.send(&layer_evals[sibling_idx].to_bytes()); // :161
        channel.send(&merkle_tree.get_authentication_path(sibling_idx)); // :162-163
    }
}

/// The same messages from ONE device launch for all layers (stark_fri_open): per layer
/// BE8(evals[idx]) || path(idx) || BE8(evals[sib]) || path(sib), fed to the channel in the reference's order.
pub fn decommit_fri_layers_proof<const M: u64>(index: usize, proof: &FRIProof<M>, channel: &mut Channel<M>) {
    let indices = [index as u64];
    let mut len = 0usize;
    ffi::check(unsafe { ffi::stark_fri_open(proof.raw(), indices.as_ptr(), 1, std::ptr::null_mut(), 0, &mut len) });
    let mut blob = vec![0u8; len];
    ffi::check(unsafe { ffi::stark_fri_open(proof.raw(), indices.as_ptr(), 1, blob.as_mut_ptr(), blob.len(), &mut len) });
    let mut off = 0usize;
    for k in 0..proof.num_layers() {
        let length = proof.layer_len(k);
        if length == 1 {
            channel.send(&blob[off..off + 8]); // :147-149
        }
        let idx = index % length;
        let sibling_idx = (idx + length / 2) % length;
        for which in [idx, sibling_idx] {
            let path_len = path_bytes(length, which);
            channel.send(&blob[off..off + 8]);
            channel.send(&blob[off + 8..off + 8 + path_len]);
            off += 8 + path_len;
        }
    }
    debug_assert_eq!(off, blob.len());
}

/// Bytes of the authentication path of leaf `idx` in a tree of `n` leaves: 32 per level that has a sibling.
fn path_bytes(n: usize, idx: usize) -> usize {
    let (mut m, mut j, mut bytes) = (n, idx, 0usize);
    while m > 1 {
        if (j ^ 1) < m {
            bytes += 32;
        }
        m = (m + 1) / 2;
        j >>= 1;
    }
    bytes
}

/// decommit_fri (reference :168-179), signature as in the reference.
pub fn decommit_fri<const M: u64>(num_queries: usize, max_index: usize, fri_layers: &[Vec<FieldElement<M>>], fri_merkles: &[MerkleTree<M>],
                                  channel: &mut Channel<M>) {
    for _ in 0..num_queries {
        let idx = channel.receive_random_int(0, max_index, true); // :176
        decommit_fri_layers(idx, fri_layers, fri_merkles, channel); // :177
    }
}

/// The same query phase against the device-resident proof: one launch per query instead of 2 x layers.
pub fn decommit_fri_proof<const M: u64>(num_queries: usize, max_index: usize, proof: &FRIProof<M>, channel: &mut Channel<M>) {
    for _ in 0..num_queries {
        let idx = channel.receive_random_int(0, max_index, true);
        decommit_fri_layers_proof(idx, proof, channel);
    }
}

// ---- several GPUs: layer 0 through the four-step NTT, hashed in leaf ranges --------------------------

/// One group of ranks, one GPU each.  Rank 0 makes the id with `MultiGpu::unique_id()` and hands the 128 bytes to the
/// other ranks by whatever the host program has (MPI, a file); every rank then calls `MultiGpu::new`.
pub struct MultiGpu<const M: u64> {
    inner: *mut ffi::stark_mg,
}

impl<const M: u64> MultiGpu<M> {
    pub fn unique_id() -> [u8; 128] {
        let mut id = [0u8; 128];
        ffi::check(unsafe { ffi::stark_mg_unique_id(id.as_mut_ptr()) });
        id
    }
    pub fn new(id: &[u8; 128], rank: u32, world: u32) -> Self {
        let mut mg: *mut ffi::stark_mg = std::ptr::null_mut();
        ffi::check(unsafe { ffi::stark_mg_create(ffi::ctx::<M>(), id.as_ptr(), rank, world, &mut mg) });
        MultiGpu { inner: mg }
    }
    pub fn rank(&self) -> u32 {
        unsafe { ffi::stark_mg_rank(self.inner) }
    }
    pub fn world(&self) -> u32 {
        unsafe { ffi::stark_mg_world(self.inner) }
    }

    /// Column-parallel LDE + commitment (BASELINE cfg4): `columns[c]` is read on rank c % world only (the others may be
    /// empty); returns the roots of all columns, identical on every rank.
    pub fn commit_columns(&self, columns: &[Vec<FieldElement<M>>], log_rows: u32, offset_in: FieldElement<M>, log_blowup: u32,
                          offset_out: FieldElement<M>) -> Vec<[u8; 32]> {
        let ptrs: Vec<*const u64> = columns.iter().map(|c| if c.is_empty() { std::ptr::null() } else { ffi::as_u64_ptr(c) }).collect();
        let mut roots = vec![0u8; 32 * columns.len()];
        ffi::check(unsafe {
            ffi::stark_mg_commit_columns(self.inner, columns.len(), ptrs.as_ptr(), log_rows, offset_in.value(), log_blowup, offset_out.value(),
                                         roots.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut())
        });
        roots.chunks_exact(32).map(|r| { let mut a = [0u8; 32]; a.copy_from_slice(r); a }).collect()
    }

    /// fri_commit + decommit_fri with the sharded layer 0 (BASELINE cfg5).  `channel` is Some on rank 0 only, where the
    /// library's own transcript object is replayed into it message by message, so that the Rust `Channel` ends in the
    /// state a single-GPU `fri_commit` + `decommit_fri` leaves behind.  transport: 0 = NCCL all-to-all, 1 = peer memory.
    pub fn fri_commit_and_decommit(&self, poly: &Polynomial<M>, log_n: u32, offset: FieldElement<M>, transport: i32, num_queries: usize,
                                   max_index: usize, channel: Option<&mut Channel<M>>) {
        let c = ffi::ctx::<M>();
        let mut coeffs: *mut ffi::stark_vec = std::ptr::null_mut();
        ffi::check(unsafe { ffi::stark_vec_upload(c, ffi::as_u64_ptr(&poly.coefficients), poly.coefficients.len(), &mut coeffs) });
        let mut lib_channel: *mut ffi::stark_channel = std::ptr::null_mut();
        if self.rank() == 0 {
            ffi::check(unsafe { ffi::stark_channel_new(M, &mut lib_channel) });
        }
        let mut proof: *mut ffi::stark_mg_fri = std::ptr::null_mut();
        ffi::check(unsafe { ffi::stark_mg_fri_commit(self.inner, coeffs, log_n, offset.value(), transport, lib_channel, &mut proof) });
        ffi::check(unsafe { ffi::stark_mg_decommit_fri(proof, num_queries, max_index, lib_channel) });
        if let Some(ch) = channel {
            replay_into(lib_channel, ch);
        }
        unsafe {
            ffi::stark_mg_fri_destroy(proof);
            ffi::stark_vec_destroy(coeffs);
            if !lib_channel.is_null() {
                ffi::stark_channel_destroy(lib_channel);
            }
        }
    }
}

impl<const M: u64> Drop for MultiGpu<M> {
    fn drop(&mut self) {
        unsafe { ffi::stark_mg_destroy(self.inner) }
    }
}

/// Copies the transcript of a library channel into a Rust `Channel` that was in the same state when the library channel
/// was created (both fresh): proof, compressed proof and state, field for field (channel.rs:14-20).
fn replay_into<const M: u64>(lib_channel: *const ffi::stark_channel, channel: &mut Channel<M>) {
    let n = unsafe { ffi::stark_channel_proof_len(lib_channel) };
    for i in 0..n {
        let mut data: *const u8 = std::ptr::null();
        let len = unsafe { ffi::stark_channel_proof_msg(lib_channel, i, &mut data) };
        channel.proof.push(unsafe { std::slice::from_raw_parts(data, len) }.to_vec());
    }
    let m = unsafe { ffi::stark_channel_compressed_len(lib_channel) };
    for i in 0..m {
        let mut data: *const u8 = std::ptr::null();
        let len = unsafe { ffi::stark_channel_compressed_msg(lib_channel, i, &mut data) };
        channel.compressed_proof.push(unsafe { std::slice::from_raw_parts(data, len) }.to_vec());
    }
    channel.state = unsafe { std::ffi::CStr::from_ptr(ffi::stark_channel_state(lib_channel)) }.to_string_lossy().into_owned();
}
