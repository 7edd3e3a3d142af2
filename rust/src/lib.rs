//! UNVERIFIED: written without a Rust toolchain (none in the build image); never compiled.
//!
//! `Lapper<T>` with `I = u32` over the C ABI of `include/lapper_b200.h`: the same method names as
//! rust-lapper 1.3.0 (`new`, `find`, `seek`, `count`, `merge_overlaps`, `cov`, `set_cov`,
//! `union_and_intersect`) plus `count_batch` / `find_batch`.  `val: T` stays on the host; the engine
//! returns positions into the sorted `intervals` vector and `perm` maps them to input positions.
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct LapperHandle {
    _private: [u8; 0],
}
#[repr(C)]
pub struct LapperResult {
    _private: [u8; 0],
}

extern "C" {
    fn lapper_last_error() -> *const c_char;
    fn lapper_build_u32(starts: *const u32, stops: *const u32, n: u64, device: c_int, out: *mut *mut LapperHandle) -> c_int;
    fn lapper_free(l: *mut LapperHandle);
    fn lapper_len(l: *const LapperHandle) -> u64;
    fn lapper_max_len(l: *const LapperHandle) -> u32;
    fn lapper_overlaps_merged(l: *const LapperHandle) -> c_int;
    fn lapper_get_intervals(l: *const LapperHandle, s: *mut u32, e: *mut u32, perm: *mut u32) -> c_int;
    fn lapper_count_batch_u32(l: *const LapperHandle, qs: *const u32, qe: *const u32, nq: u64, out: *mut u64) -> c_int;
    fn lapper_find_batch_u32(l: *const LapperHandle, qs: *const u32, qe: *const u32, nq: u64, out: *mut *mut LapperResult) -> c_int;
    fn lapper_result_total(r: *const LapperResult) -> u64;
    fn lapper_result_copy(r: *const LapperResult, offsets: *mut u64, indices: *mut u32) -> c_int;
    fn lapper_result_free(r: *mut LapperResult);
    fn lapper_seek_cursor(l: *const LapperHandle, start: u32, cursor: *mut u64) -> c_int;
    fn lapper_merge_overlaps(l: *mut LapperHandle) -> c_int;
    fn lapper_cov(l: *const LapperHandle, out: *mut u32) -> c_int;
    fn lapper_set_cov(l: *mut LapperHandle, out: *mut u32) -> c_int;
    fn lapper_union_and_intersect(a: *const LapperHandle, b: *const LapperHandle, u: *mut u32, i: *mut u32) -> c_int;
    fn lapper_insert_u32(l: *mut LapperHandle, start: u32, stop: u32, input_position_out: *mut u64) -> c_int;
    fn lapper_depth(l: *const LapperHandle, n: *mut u64, starts: *mut *mut u32, stops: *mut *mut u32, depth: *mut *mut u32) -> c_int;
    fn lapper_host_free(p: *mut c_void);
    fn lapper_find_batch_u32_host(l: *const LapperHandle, qs: *const u32, qe: *const u32, nq: u64, offsets: *mut u64,
                                  indices: *mut u32, capacity: u64, total: *mut u64) -> c_int;
    fn lapper_get_cached_state(l: *const LapperHandle, cov_is_some: *mut c_int, cov: *mut u32, overlaps_merged: *mut c_int) -> c_int;
    fn lapper_restore_cached_state(l: *mut LapperHandle, cov_is_some: c_int, cov: u32, overlaps_merged: c_int) -> c_int;
    fn lapper_tune_query_length(l: *mut LapperHandle, typical_query_len: u32) -> c_int;
}

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(lapper_last_error()) }.to_string_lossy().into_owned();
        panic!("liblapper_b200 error {}: {}", rc, msg); // the reference's only failure mode is a panic too
    }
}

#[derive(Debug, Clone, Eq)]
pub struct Interval<T: Eq + Clone + Send + Sync> {
    pub start: u32,
    pub stop: u32,
    pub val: T,
}
impl<T: Eq + Clone + Send + Sync> PartialEq for Interval<T> {
    fn eq(&self, o: &Self) -> bool {
        self.start == o.start && self.stop == o.stop
    }
}

pub struct Lapper<T: Eq + Clone + Send + Sync> {
    pub intervals: Vec<Interval<T>>,
    input: Vec<Interval<T>>,
    h: *mut LapperHandle,
    _pin: *mut c_void,
}
unsafe impl<T: Eq + Clone + Send + Sync> Send for Lapper<T> {}
unsafe impl<T: Eq + Clone + Send + Sync> Sync for Lapper<T> {} // const calls on one handle are thread-safe

impl<T: Eq + Clone + Send + Sync> Lapper<T> {
    pub fn new(intervals: Vec<Interval<T>>) -> Self {
        let s: Vec<u32> = intervals.iter().map(|x| x.start).collect();
        let e: Vec<u32> = intervals.iter().map(|x| x.stop).collect();
        let mut h = std::ptr::null_mut();
        check(unsafe { lapper_build_u32(s.as_ptr(), e.as_ptr(), s.len() as u64, 0, &mut h) });
        let mut l = Lapper { intervals: vec![], input: intervals, h, _pin: std::ptr::null_mut() };
        l.refresh();
        l
    }
    fn refresh(&mut self) {
        let n = unsafe { lapper_len(self.h) } as usize;
        let (mut s, mut e, mut p) = (vec![0u32; n], vec![0u32; n], vec![0u32; n]);
        check(unsafe { lapper_get_intervals(self.h, s.as_mut_ptr(), e.as_mut_ptr(), p.as_mut_ptr()) });
        self.intervals = (0..n).map(|i| Interval { start: s[i], stop: e[i], val: self.input[p[i] as usize].val.clone() }).collect();
    }
    /// Not in the reference: announce the typical query length so that random-order `count` stays on its
    /// one-sector path for reads longer than 150 bp (results never change).
    pub fn tune_query_length(&mut self, typical_query_len: u32) { check(unsafe { lapper_tune_query_length(self.h, typical_query_len) }) }
    pub fn len(&self) -> usize { self.intervals.len() }
    pub fn is_empty(&self) -> bool { self.intervals.is_empty() }
    pub fn max_len(&self) -> u32 { unsafe { lapper_max_len(self.h) } }
    pub fn overlaps_merged(&self) -> bool { unsafe { lapper_overlaps_merged(self.h) != 0 } }
    pub fn iter(&self) -> std::slice::Iter<'_, Interval<T>> { self.intervals.iter() }

    pub fn count_batch(&self, qs: &[u32], qe: &[u32]) -> Vec<u64> {
        assert_eq!(qs.len(), qe.len());
        let mut out = vec![0u64; qs.len()];
        check(unsafe { lapper_count_batch_u32(self.h, qs.as_ptr(), qe.as_ptr(), qs.len() as u64, out.as_mut_ptr()) });
        out
    }
    /// CSR: `indices[offsets[j]..offsets[j+1]]` are the positions `find(qs[j], qe[j])` yields, in order.
    pub fn find_batch(&self, qs: &[u32], qe: &[u32]) -> (Vec<u64>, Vec<u32>) {
        assert_eq!(qs.len(), qe.len());
        let mut r = std::ptr::null_mut();
        check(unsafe { lapper_find_batch_u32(self.h, qs.as_ptr(), qe.as_ptr(), qs.len() as u64, &mut r) });
        let total = unsafe { lapper_result_total(r) } as usize;
        let (mut off, mut idx) = (vec![0u64; qs.len() + 1], vec![0u32; total]);
        let rc = unsafe { lapper_result_copy(r, off.as_mut_ptr(), idx.as_mut_ptr()) };
        unsafe { lapper_result_free(r) };
        check(rc);
        (off, idx)
    }
    pub fn count(&self, start: u32, stop: u32) -> usize { self.count_batch(&[start], &[stop])[0] as usize }
    pub fn find(&self, start: u32, stop: u32) -> impl Iterator<Item = &Interval<T>> {
        let (_, idx) = self.find_batch(&[start], &[stop]);
        idx.into_iter().map(move |i| &self.intervals[i as usize])
    }
    pub fn seek<'a>(&'a self, start: u32, stop: u32, cursor: &mut usize) -> impl Iterator<Item = &'a Interval<T>> {
        let mut c = *cursor as u64;
        check(unsafe { lapper_seek_cursor(self.h, start, &mut c) });
        *cursor = c as usize;
        let (_, idx) = self.find_batch(&[start], &[stop]);
        idx.into_iter().filter(move |&i| i as u64 >= c).map(move |i| &self.intervals[i as usize])
    }
    pub fn merge_overlaps(&mut self) {
        check(unsafe { lapper_merge_overlaps(self.h) });
        self.refresh();
    }
    pub fn cov(&self) -> u32 { let mut c = 0; check(unsafe { lapper_cov(self.h, &mut c) }); c }
    pub fn set_cov(&mut self) -> u32 { let mut c = 0; check(unsafe { lapper_set_cov(self.h, &mut c) }); c }
    pub fn union_and_intersect(&self, other: &Self) -> (u32, u32) {
        let (mut u, mut i) = (0, 0);
        check(unsafe { lapper_union_and_intersect(self.h, other.h, &mut u, &mut i) });
        (u, i)
    }
    pub fn intersect(&self, other: &Self) -> u32 { self.union_and_intersect(other).1 }
    pub fn union(&self, other: &Self) -> u32 { self.union_and_intersect(other).0 }

    /// src/lib.rs:260-273.  O(n) like the reference; the engine hands back the input position it gave the interval.
    pub fn insert(&mut self, elem: Interval<T>) {
        let mut pos = 0u64;
        check(unsafe { lapper_insert_u32(self.h, elem.start, elem.stop, &mut pos) });
        debug_assert_eq!(pos as usize, self.input.len());
        self.input.push(elem);
        self.refresh();
    }

    /// src/lib.rs:576-598: maximal runs of equal coverage depth, as `Interval<u32>` with `val` = depth.
    /// Panics on an empty lapper like the reference (src/lib.rs:744).
    pub fn depth(&self) -> impl Iterator<Item = Interval<u32>> {
        let (mut n, mut s, mut e, mut d) = (0u64, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut());
        check(unsafe { lapper_depth(self.h, &mut n, &mut s, &mut e, &mut d) });
        let n = n as usize;
        let out: Vec<Interval<u32>> = (0..n)
            .map(|i| unsafe { Interval { start: *s.add(i), stop: *e.add(i), val: *d.add(i) } })
            .collect();
        unsafe {
            lapper_host_free(s as *mut c_void);
            lapper_host_free(e as *mut c_void);
            lapper_host_free(d as *mut c_void);
        }
        out.into_iter()
    }

    /// find_batch straight into caller-owned buffers through the pipelined host entry point; returns the total.
    /// `Err(total)` when `indices` is too small (offsets are complete; call again with `total` slots).
    pub fn find_batch_into(&self, qs: &[u32], qe: &[u32], offsets: &mut [u64], indices: &mut [u32]) -> Result<u64, u64> {
        assert!(qs.len() == qe.len() && offsets.len() > qs.len());
        let mut total = 0u64;
        let rc = unsafe {
            lapper_find_batch_u32_host(self.h, qs.as_ptr(), qe.as_ptr(), qs.len() as u64, offsets.as_mut_ptr(),
                                       indices.as_mut_ptr(), indices.len() as u64, &mut total)
        };
        match rc {
            0 => Ok(total),
            6 => Err(total), // LAPPER_ERR_CAPACITY
            _ => { check(rc); unreachable!() }
        }
    }

    /// The `with_serde` feature (src/lib.rs:85-86): serialise `intervals` + the cached `cov` / `overlaps_merged`
    /// (lapper_get_cached_state), rebuild the index on load and restore the cache (lapper_restore_cached_state).
    /// rust_lapper_b200/bincode_io.py is the exercised implementation of that layout.
    pub fn cached_state(&self) -> (Option<u32>, bool) {
        let (mut some, mut cov, mut merged) = (0, 0u32, 0);
        check(unsafe { lapper_get_cached_state(self.h, &mut some, &mut cov, &mut merged) });
        (if some != 0 { Some(cov) } else { None }, merged != 0)
    }
    pub fn restore_cached_state(&mut self, cov: Option<u32>, overlaps_merged: bool) {
        check(unsafe { lapper_restore_cached_state(self.h, cov.is_some() as c_int, cov.unwrap_or(0), overlaps_merged as c_int) });
    }
}
/// src/lib.rs:785-849: `IntoIterator` for `Lapper` and `&Lapper` walk the sorted `intervals` vector.
impl<T: Eq + Clone + Send + Sync> IntoIterator for Lapper<T> {
    type Item = Interval<T>;
    type IntoIter = std::vec::IntoIter<Interval<T>>;
    fn into_iter(mut self) -> Self::IntoIter { std::mem::take(&mut self.intervals).into_iter() }
}
impl<'a, T: Eq + Clone + Send + Sync> IntoIterator for &'a Lapper<T> {
    type Item = &'a Interval<T>;
    type IntoIter = std::slice::Iter<'a, Interval<T>>;
    fn into_iter(self) -> Self::IntoIter { self.intervals.iter() }
}
impl<T: Eq + Clone + Send + Sync> Drop for Lapper<T> {
    fn drop(&mut self) { unsafe { lapper_free(self.h) } }
}

/// `Lapper<usize, T>` -- the coordinate type of the reference's README examples -- over `lapper64_*`
/// (include/lapper_b200.h).  Sets whose coordinates fit 32 bits are answered by the u32 engine inside the library
/// (narrowing, csrc/engine64.cu); nothing changes for the caller.  UNVERIFIED like the rest of this file.
pub mod wide {
    use super::{check, lapper_result_copy, lapper_result_free, lapper_result_total, LapperResult};
    use std::os::raw::c_int;

    #[repr(C)]
    pub struct Lapper64Handle {
        _private: [u8; 0],
    }
    extern "C" {
        fn lapper64_build(starts: *const u64, stops: *const u64, n: u64, device: c_int, out: *mut *mut Lapper64Handle) -> c_int;
        fn lapper64_free(l: *mut Lapper64Handle);
        fn lapper64_len(l: *const Lapper64Handle) -> u64;
        fn lapper64_max_len(l: *const Lapper64Handle) -> u64;
        fn lapper64_get_intervals(l: *const Lapper64Handle, s: *mut u64, e: *mut u64, perm: *mut u32) -> c_int;
        fn lapper64_count_batch(l: *const Lapper64Handle, qs: *const u64, qe: *const u64, nq: u64, out: *mut u64) -> c_int;
        fn lapper64_find_batch(l: *const Lapper64Handle, qs: *const u64, qe: *const u64, nq: u64, out: *mut *mut LapperResult) -> c_int;
        fn lapper64_merge_overlaps(l: *mut Lapper64Handle) -> c_int;
        fn lapper64_cov(l: *const Lapper64Handle, out: *mut u64) -> c_int;
        fn lapper64_union_and_intersect(a: *const Lapper64Handle, b: *const Lapper64Handle, u: *mut u64, i: *mut u64) -> c_int;
        fn lapper64_insert(l: *mut Lapper64Handle, start: u64, stop: u64, input_position_out: *mut u64) -> c_int;
        fn lapper64_seek_cursor(l: *const Lapper64Handle, start: u64, cursor: *mut u64) -> c_int;
        fn lapper64_set_cov(l: *mut Lapper64Handle, out: *mut u64) -> c_int;
        fn lapper64_get_cached_state(l: *const Lapper64Handle, cov_is_some: *mut c_int, cov: *mut u64, overlaps_merged: *mut c_int) -> c_int;
        fn lapper64_restore_cached_state(l: *mut Lapper64Handle, cov_is_some: c_int, cov: u64, overlaps_merged: c_int) -> c_int;
    }

    #[derive(Debug, Clone)]
    pub struct Interval<T: Eq + Clone + Send + Sync> {
        pub start: usize,
        pub stop: usize,
        pub val: T,
    }

    pub struct Lapper<T: Eq + Clone + Send + Sync> {
        pub intervals: Vec<Interval<T>>,
        input: Vec<Interval<T>>,
        h: *mut Lapper64Handle,
    }
    unsafe impl<T: Eq + Clone + Send + Sync> Send for Lapper<T> {}
    unsafe impl<T: Eq + Clone + Send + Sync> Sync for Lapper<T> {}

    impl<T: Eq + Clone + Send + Sync> Lapper<T> {
        pub fn new(intervals: Vec<Interval<T>>) -> Self {
            let s: Vec<u64> = intervals.iter().map(|x| x.start as u64).collect();
            let e: Vec<u64> = intervals.iter().map(|x| x.stop as u64).collect();
            let mut h = std::ptr::null_mut();
            check(unsafe { lapper64_build(s.as_ptr(), e.as_ptr(), s.len() as u64, 0, &mut h) });
            let mut l = Lapper { intervals: vec![], input: intervals, h };
            l.refresh();
            l
        }
        fn refresh(&mut self) {
            let n = unsafe { lapper64_len(self.h) } as usize;
            let (mut s, mut e, mut p) = (vec![0u64; n], vec![0u64; n], vec![0u32; n]);
            check(unsafe { lapper64_get_intervals(self.h, s.as_mut_ptr(), e.as_mut_ptr(), p.as_mut_ptr()) });
            self.intervals = (0..n)
                .map(|i| Interval { start: s[i] as usize, stop: e[i] as usize, val: self.input[p[i] as usize].val.clone() })
                .collect();
        }
        pub fn len(&self) -> usize { self.intervals.len() }
        pub fn is_empty(&self) -> bool { self.intervals.is_empty() }
        pub fn max_len(&self) -> usize { unsafe { lapper64_max_len(self.h) as usize } }
        pub fn count_batch(&self, qs: &[u64], qe: &[u64]) -> Vec<u64> {
            assert_eq!(qs.len(), qe.len());
            let mut out = vec![0u64; qs.len()];
            check(unsafe { lapper64_count_batch(self.h, qs.as_ptr(), qe.as_ptr(), qs.len() as u64, out.as_mut_ptr()) });
            out
        }
        pub fn find_batch(&self, qs: &[u64], qe: &[u64]) -> (Vec<u64>, Vec<u32>) {
            assert_eq!(qs.len(), qe.len());
            let mut r = std::ptr::null_mut();
            check(unsafe { lapper64_find_batch(self.h, qs.as_ptr(), qe.as_ptr(), qs.len() as u64, &mut r) });
            let total = unsafe { lapper_result_total(r) } as usize;
            let (mut off, mut idx) = (vec![0u64; qs.len() + 1], vec![0u32; total]);
            let rc = unsafe { lapper_result_copy(r, off.as_mut_ptr(), idx.as_mut_ptr()) };
            unsafe { lapper_result_free(r) };
            check(rc);
            (off, idx)
        }
        pub fn count(&self, start: usize, stop: usize) -> usize { self.count_batch(&[start as u64], &[stop as u64])[0] as usize }
        pub fn find(&self, start: usize, stop: usize) -> impl Iterator<Item = &Interval<T>> {
            let (_, idx) = self.find_batch(&[start as u64], &[stop as u64]);
            idx.into_iter().map(move |i| &self.intervals[i as usize])
        }
        pub fn seek<'a>(&'a self, start: usize, stop: usize, cursor: &mut usize) -> impl Iterator<Item = &'a Interval<T>> {
            let mut c = *cursor as u64;
            check(unsafe { lapper64_seek_cursor(self.h, start as u64, &mut c) });
            *cursor = c as usize;
            let (_, idx) = self.find_batch(&[start as u64], &[stop as u64]);
            idx.into_iter().filter(move |&i| i as u64 >= c).map(move |i| &self.intervals[i as usize])
        }
        pub fn insert(&mut self, elem: Interval<T>) {
            let mut pos = 0u64;
            check(unsafe { lapper64_insert(self.h, elem.start as u64, elem.stop as u64, &mut pos) });
            debug_assert_eq!(pos as usize, self.input.len());
            self.input.push(elem);
            self.refresh();
        }
        pub fn merge_overlaps(&mut self) {
            check(unsafe { lapper64_merge_overlaps(self.h) });
            self.refresh();
        }
        pub fn cov(&self) -> usize { let mut c = 0u64; check(unsafe { lapper64_cov(self.h, &mut c) }); c as usize }
        pub fn union_and_intersect(&self, other: &Self) -> (usize, usize) {
            let (mut u, mut i) = (0u64, 0u64);
            check(unsafe { lapper64_union_and_intersect(self.h, other.h, &mut u, &mut i) });
            (u as usize, i as usize)
        }
        pub fn set_cov(&mut self) -> usize { let mut c = 0u64; check(unsafe { lapper64_set_cov(self.h, &mut c) }); c as usize }
        /// `with_serde` for `Lapper<usize, T>`: see `cached_state` / `restore_cached_state` of the u32 form above;
        /// include/lapper_b200.hpp (`to_bincode` / `from_bincode`) is the exercised implementation of the layout.
        pub fn cached_state(&self) -> (Option<usize>, bool) {
            let (mut some, mut cov, mut merged) = (0, 0u64, 0);
            check(unsafe { lapper64_get_cached_state(self.h, &mut some, &mut cov, &mut merged) });
            (if some != 0 { Some(cov as usize) } else { None }, merged != 0)
        }
        pub fn restore_cached_state(&mut self, cov: Option<usize>, overlaps_merged: bool) {
            check(unsafe { lapper64_restore_cached_state(self.h, cov.is_some() as c_int, cov.unwrap_or(0) as u64, overlaps_merged as c_int) });
        }
    }
    impl<T: Eq + Clone + Send + Sync> Drop for Lapper<T> {
        fn drop(&mut self) { unsafe { lapper64_free(self.h) } }
    }
}
