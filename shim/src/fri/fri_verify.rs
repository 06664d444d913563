//! src/fri/fri_verify.rs — verify_fri (reference src/fri/fri_verify.rs:12-177) completed: the Merkle check the draft
//! calls and never defines (:109,137) and the fold-consistency check it comments out (:153-170), host side.
use crate::channel::Channel;
use crate::ffi;
use crate::fields::FieldElement;

/// Flattens `channel.proof` the way the verifier reads it: u32-LE length || bytes per message.
pub fn flatten_proof<const M: u64>(channel: &Channel<M>) -> Vec<u8> {
    let mut out = Vec::with_capacity(channel.proof_size() + 4 * channel.proof.len());
    for m in channel.proof.iter() {
        out.extend_from_slice(&(m.len() as u32).to_le_bytes());
        out.extend_from_slice(m);
    }
    out
}

/// Replays a proof made by `fri_commit` + `decommit_fri` against a fresh channel.  `expected_num_layers` is the
/// draft's parameter (:15): the verifier's bound on the number of layers, i.e. the committed polynomial is claimed to
/// have at most 2^(expected_num_layers - 1) coefficients; a proof with more folds is rejected.
/// Returns Ok(()) or the first failed check.
pub fn verify_fri<const M: u64>(proof: &[u8], log_domain: u32, offset: FieldElement<M>, num_queries: usize, max_index: usize,
                                expected_num_layers: usize) -> Result<(), String> {
    assert!(expected_num_layers >= 1, "a FRI proof has at least one layer");
    let generator = unsafe { ffi::stark_ctx_generator(ffi::ctx::<M>()) };
    let mut ok = 0;
    let mut reason = [0 as std::os::raw::c_char; 160];
    ffi::check(unsafe {
        ffi::stark_fri_verify(proof.as_ptr(), proof.len(), M, generator, log_domain, offset.value(), num_queries, max_index,
                              (expected_num_layers - 1) as u32, &mut ok, reason.as_mut_ptr())
    });
    if ok == 1 {
        Ok(())
    } else {
        Err(unsafe { std::ffi::CStr::from_ptr(reason.as_ptr()) }.to_string_lossy().into_owned())
    }
}
