//! src/merkle/mod.rs — MerkleTree over field elements, hashed on the device.
//!
//! Same public surface as the reference (`new`, `root`, reference src/merkle/mod.rs:10-26) plus the two methods the
//! crate calls and never defines: `get_authentication_path` (fri_commit.rs:157,162) and `validate` (fri_verify.rs:109,137).
//! Tree rule (rs_merkle 1.4.2 `from_leaves` / `root_hex`): leaf = SHA-256(value.to_be_bytes()), parent = SHA-256(l || r),
//! a node without a right sibling is promoted unchanged, root as lowercase hex.
use std::os::raw::c_char;

use crate::ffi;
use crate::fields::FieldElement;

pub struct MerkleTree<const MODULUS: u64> {
    inner: *mut ffi::stark_tree,
    owned: bool, // false: borrowed from a FRIProof (the proof owns the device memory)
}

// the tree lives in HBM behind a context that serialises its callers
unsafe impl<const MODULUS: u64> Send for MerkleTree<MODULUS> {}
unsafe impl<const MODULUS: u64> Sync for MerkleTree<MODULUS> {}

impl<const MODULUS: u64> MerkleTree<MODULUS> {
    pub fn new(data: Vec<FieldElement<MODULUS>>) -> Self {
        // reference :10-22.  An empty vector is an error here where the reference panics later, in root() (:25).
        let mut tree: *mut ffi::stark_tree = std::ptr::null_mut();
        ffi::check(unsafe { ffi::stark_merkle_commit(ffi::ctx::<MODULUS>(), ffi::as_u64_ptr(&data), data.len(), &mut tree) });
        MerkleTree { inner: tree, owned: true }
    }

    /// A tree that already lives on the device (a FRI layer's): not destroyed on drop.
    pub(crate) fn borrowed(tree: *const ffi::stark_tree) -> Self {
        MerkleTree { inner: tree as *mut ffi::stark_tree, owned: false }
    }

    pub fn root(&self) -> String {
        // reference :24-26
        let mut buf = [0 as c_char; 65];
        ffi::check(unsafe { ffi::stark_merkle_root_hex(self.inner, buf.as_mut_ptr()) });
        unsafe { std::ffi::CStr::from_ptr(buf.as_ptr()) }.to_str().expect("hex root").to_owned()
    }

    pub fn root_bytes(&self) -> [u8; 32] {
        let mut out = [0u8; 32];
        ffi::check(unsafe { ffi::stark_merkle_root(self.inner, out.as_mut_ptr()) });
        out
    }

    pub fn num_leaves(&self) -> usize {
        unsafe { ffi::stark_merkle_num_leaves(self.inner) }
    }

    /// Sibling digests bottom -> top, 32 bytes each (rs_merkle `MerkleProof::to_bytes` for one leaf).
    pub fn get_authentication_path(&self, idx: usize) -> Vec<u8> {
        let depth = unsafe { ffi::stark_merkle_depth(self.inner) };
        let mut path = vec![0u8; 32 * (depth + 1)];
        let mut len = 0usize;
        ffi::check(unsafe { ffi::stark_merkle_open(self.inner, idx, path.as_mut_ptr(), path.len(), &mut len) });
        path.truncate(len);
        path
    }

    /// rs_merkle `MerkleProof::verify` for one leaf (host side).
    pub fn validate(root: &[u8; 32], num_leaves: usize, idx: usize, value: FieldElement<MODULUS>, path: &[u8]) -> bool {
        let mut ok = 0;
        ffi::check(unsafe { ffi::stark_merkle_verify(root.as_ptr(), num_leaves, idx, value.value(), path.as_ptr(), path.len(), &mut ok) });
        ok == 1
    }

    pub(crate) fn raw(&self) -> *const ffi::stark_tree {
        self.inner
    }
}

impl<const MODULUS: u64> Drop for MerkleTree<MODULUS> {
    fn drop(&mut self) {
        if self.owned && !self.inner.is_null() {
            unsafe { ffi::stark_tree_destroy(self.inner) }
        }
    }
}
