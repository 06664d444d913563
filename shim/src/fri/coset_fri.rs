//! src/fri/coset_fri.rs — coset domains D = { offset * omega^i } (reference src/fri/coset_fri.rs:22-52).
use crate::ffi;
use crate::fields::element::FieldElement;

#[derive(Clone, Debug)]
pub struct CosetFri<const M: u64> {
    pub offset: FieldElement<M>,
    pub omega: FieldElement<M>,
    pub domain_size: usize,
}

impl<const M: u64> CosetFri<M> {
    pub fn new(offset: FieldElement<M>, omega: FieldElement<M>, domain_size: usize) -> Self {
        Self { offset, omega, domain_size }
    }

    /// The coset whose generator is the library's root of unity of that order (the one every device transform uses).
    pub fn with_library_root(offset: FieldElement<M>, domain_size: usize) -> Self {
        assert!(domain_size.is_power_of_two(), "domain size must be a power of two");
        let w = unsafe { ffi::stark_ctx_root_of_unity(ffi::ctx::<M>(), domain_size.trailing_zeros()) };
        assert!(w != 0, "2^k does not divide p - 1: no root of unity of that order");
        Self { offset, omega: FieldElement::new(w), domain_size }
    }

    /// D = { offset * omega^i : i in [0, domain_size) }, natural order (reference :32-36).
    pub fn generate_coset_domain(&self) -> Vec<FieldElement<M>> {
        let c = ffi::ctx::<M>();
        let log_n = self.domain_size.trailing_zeros();
        let on_device = self.domain_size.is_power_of_two()
            && log_n <= unsafe { ffi::stark_ctx_two_adicity(c) }
            && self.omega.value() == unsafe { ffi::stark_ctx_root_of_unity(c, log_n) }
            && self.offset != FieldElement::<M>::zero();
        if on_device {
            let mut out = vec![FieldElement::<M>::zero(); self.domain_size];
            ffi::check(unsafe { ffi::stark_coset_domain(c, log_n, self.offset.value(), ffi::as_u64_mut_ptr(&mut out)) });
            out
        } else {
            // any other generator: the reference's loop, one multiplication per element instead of a pow
            let mut out = Vec::with_capacity(self.domain_size);
            let mut x = self.offset;
            for _ in 0..self.domain_size {
                out.push(x);
                x = x * self.omega;
            }
            out
        }
    }

    /// Squares every element (reference :40-51, "square all" as written there).
    pub fn next_coset_domain(&self, current_domain: &[FieldElement<M>]) -> Vec<FieldElement<M>> {
        current_domain.iter().map(|d| d.square()).collect()
    }
}
