//! src/fri/mod.rs — the FRI module, re-enabled in lib.rs (it was commented out: the draft did not compile).
pub mod coset_fri;
pub mod fri_commit;
pub mod fri_verify;

pub use coset_fri::*;
pub use fri_commit::*;
pub use fri_verify::*;
