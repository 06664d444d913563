//! src/polynomial/gpu.rs — the three places where the reference's polynomial code is data parallel, routed to the
//! device; everything else in ops.rs / interpolation.rs is untouched host Rust.
//!
//!   evaluate over a domain   ops.rs:76-83 mapped at fri_commit.rs:78            -> coset NTT
//!   interpolate              ops.rs:239-241 -> interpolation.rs:121-152          -> coset iNTT when xs is a coset
//!   division by a vanishing  ops.rs:141-191, impl Div :412-421                   -> point-wise quotient with a batched,
//!   polynomial                                                                      zero-safe inverse, then an iNTT
//!
//! Canonical results are unique, so each of these returns exactly what the host algorithm returns.
use crate::ffi;
use crate::fields::FieldElement;
use crate::polynomial::interpolation::interpolate_lagrange_polynomials;
use crate::polynomial::Polynomial;

impl<const M: u64> Polynomial<M> {
    /// `domain.iter().map(|x| self.evaluate(*x)).collect()` (fri_commit.rs:78).  Cosets of power-of-two order go to the
    /// device; any other domain keeps the Horner loop.
    pub fn evaluate_domain(&self, domain: &[FieldElement<M>]) -> Vec<FieldElement<M>> {
        match ffi::as_coset(domain) {
            Some((log_n, offset)) if self.coefficients.len() <= domain.len() => {
                let mut out = vec![FieldElement::<M>::zero(); domain.len()];
                ffi::check(unsafe {
                    ffi::stark_coset_evaluate(ffi::ctx::<M>(), ffi::as_u64_ptr(&self.coefficients), self.coefficients.len(), log_n, offset,
                                              ffi::as_u64_mut_ptr(&mut out))
                });
                out
            }
            _ => domain.iter().map(|&x| self.evaluate(x)).collect(),
        }
    }

    /// `Polynomial::interpolate` (ops.rs:239-241).  When `xs` is a coset offset * <w_n> the unique interpolant is one
    /// inverse transform; `Polynomial::new` trims it like the Lagrange path does (ops.rs:19-37).
    pub fn interpolate_fast(xs: &[FieldElement<M>], ys: &[FieldElement<M>]) -> Self {
        assert_eq!(xs.len(), ys.len(), "interpolate: xs and ys differ in length");
        match ffi::as_coset(xs) {
            Some((log_n, offset)) => {
                let mut coeffs = vec![FieldElement::<M>::zero(); xs.len()];
                ffi::check(unsafe {
                    ffi::stark_coset_interpolate(ffi::ctx::<M>(), ffi::as_u64_ptr(ys), log_n, offset, ffi::as_u64_mut_ptr(&mut coeffs))
                });
                Polynomial::new(coeffs)
            }
            None => interpolate_lagrange_polynomials(xs, ys),
        }
    }

    /// Exact division by the vanishing polynomial Z(x) = x^t - 1 of a trace domain of size t (a power of two), the
    /// division a STARK composition performs.  Evaluates numerator and Z on a coset that avoids Z's roots, divides
    /// point-wise (one batched inverse for the whole domain) and interpolates the quotient back.  Panics like `impl Div`
    /// (ops.rs:416-418) when the remainder is not zero -- detected as a quotient of too high a degree.
    pub fn div_by_vanishing(&self, t: usize) -> Self {
        assert!(t.is_power_of_two(), "vanishing polynomial of a power-of-two domain");
        if self.is_zero() {
            return Polynomial::zero();
        }
        let c = ffi::ctx::<M>();
        let deg = self.degree as usize;
        assert!(deg >= t, "Polynomial division had a non-zero remainder");
        let n = (deg + 1).next_power_of_two().max(2 * t);
        let log_n = n.trailing_zeros();
        let offset = unsafe { ffi::stark_ctx_generator(c) }; // a generator of F_p^*: its coset misses every subgroup
        let mut num = vec![FieldElement::<M>::zero(); n];
        ffi::check(unsafe {
            ffi::stark_coset_evaluate(c, ffi::as_u64_ptr(&self.coefficients), self.coefficients.len(), log_n, offset, ffi::as_u64_mut_ptr(&mut num))
        });
        // Z on the coset: (offset * w^i)^t - 1 takes n/t distinct values
        let w = FieldElement::<M>::new(unsafe { ffi::stark_ctx_root_of_unity(c, log_n) });
        let (off_t, w_t) = (FieldElement::<M>::new(offset).pow(t as u64), w.pow(t as u64));
        let mut den = Vec::with_capacity(n);
        let mut cur = off_t;
        for i in 0..n {
            if i > 0 && i % (n / t) == 0 {
                cur = off_t; // w^(t * n/t) = 1: the values repeat with period n/t
            }
            den.push(cur - FieldElement::<M>::one());
            cur = cur * w_t;
        }
        let mut quot = vec![FieldElement::<M>::zero(); n];
        ffi::check(unsafe { ffi::stark_quotient_pointwise(c, ffi::as_u64_ptr(&num), ffi::as_u64_ptr(&den), n, ffi::as_u64_mut_ptr(&mut quot)) });
        let mut coeffs = vec![FieldElement::<M>::zero(); n];
        ffi::check(unsafe { ffi::stark_coset_interpolate(c, ffi::as_u64_ptr(&quot), log_n, offset, ffi::as_u64_mut_ptr(&mut coeffs)) });
        let q = Polynomial::new(coeffs);
        if q.degree > self.degree - t as isize {
            panic!("Polynomial division had a non-zero remainder");
        }
        q
    }

    /// `a.iter().map(|x| x.inverse())` with one exponentiation per 4096 elements (element.rs:54-57; inverse(0) == 0).
    pub fn batch_inverse(values: &mut [FieldElement<M>]) {
        ffi::check(unsafe { ffi::stark_batch_inverse(ffi::ctx::<M>(), ffi::as_u64_mut_ptr(values), values.len()) });
    }
}
