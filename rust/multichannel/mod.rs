//! Drop-in content for yagi's empty `src/multichannel/mod.rs` (declared at src/lib.rs:27-28):
//! `FirPfbCh2<Complex32, f32>` over libyagi_b200.  Follows the object protocol of the existing
//! filter structs (`new*` -> `Result<Self>`, `reset`, `execute`, `execute_block`, getters, `Clone`),
//! e.g. src/filter/fir/firdecim.rs:38-57,124-126,179-205.
//!
//! SOURCE ONLY: not compiled in the development image (no Rust toolchain there).  The C ABI it
//! calls is what tests/ exercise through the Python mirror.
use crate::error::{Error, Result};
use num_complex::Complex32;
use std::ffi::CStr;
use yagi_b200_sys as sys;

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum FirPfbChType {
    Analyzer,
    Synthesizer,
}

fn check(status: i32) -> Result<()> {
    if status == sys::YG_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(sys::yg_last_error()) }.to_string_lossy().into_owned();
    Err(match status {
        sys::YG_ECONFIG => Error::Config(msg),
        sys::YG_EVALUE => Error::Value(msg),
        sys::YG_ERANGE => Error::Range(msg),
        sys::YG_EMODE => Error::Mode(msg),
        sys::YG_ENOCONV => Error::NoConvergence(msg),
        _ => Error::Internal(msg),
    })
}

/// 2x oversampled polyphase filterbank channelizer (upstream firpfbch2_crcf), GPU-backed.
#[derive(Debug)]
pub struct FirPfbCh2 {
    q: sys::yg_firpfbch2_crcf,
    type_: FirPfbChType,
    num_channels: usize,
    m: usize,
}

unsafe impl Send for FirPfbCh2 {}

impl FirPfbCh2 {
    pub fn new(type_: FirPfbChType, num_channels: usize, m: usize, h: &[f32]) -> Result<Self> {
        let mut q = std::ptr::null_mut();
        check(unsafe {
            sys::yg_firpfbch2_crcf_create(type_ as i32, num_channels as u32, m as u32, h.as_ptr(), h.len(), &mut q)
        })?;
        Ok(Self { q, type_, num_channels, m })
    }

    pub fn new_kaiser(type_: FirPfbChType, num_channels: usize, m: usize, as_: f32) -> Result<Self> {
        let mut q = std::ptr::null_mut();
        check(unsafe { sys::yg_firpfbch2_crcf_create_kaiser(type_ as i32, num_channels as u32, m as u32, as_, &mut q) })?;
        Ok(Self { q, type_, num_channels, m })
    }

    pub fn reset(&mut self) {
        let _ = unsafe { sys::yg_firpfbch2_crcf_reset(self.q) };
    }

    pub fn get_type(&self) -> FirPfbChType {
        self.type_
    }
    pub fn get_num_channels(&self) -> usize {
        self.num_channels
    }
    pub fn get_m(&self) -> usize {
        self.m
    }

    fn io(&self) -> (usize, usize) {
        match self.type_ {
            FirPfbChType::Analyzer => (self.num_channels / 2, self.num_channels),
            FirPfbChType::Synthesizer => (self.num_channels, self.num_channels / 2),
        }
    }

    /// One frame: analyzer M/2 -> M, synthesizer M -> M/2.
    pub fn execute(&mut self, x: &[Complex32], y: &mut [Complex32]) -> Result<()> {
        self.execute_block(x, 1, y)
    }

    /// `n` consecutive frames (host slices; copies happen inside the call).
    pub fn execute_block(&mut self, x: &[Complex32], n: usize, y: &mut [Complex32]) -> Result<()> {
        let (nin, nout) = self.io();
        if x.len() != n * nin || y.len() != n * nout {
            return Err(Error::Config("input/output block lengths do not match the frame count".into()));
        }
        // Complex32 is #[repr(C)] { re: f32, im: f32 } == yg_cf32
        check(unsafe {
            sys::yg_firpfbch2_crcf_execute_block(self.q, x.as_ptr() as *const sys::yg_cf32, n, y.as_mut_ptr() as *mut sys::yg_cf32)
        })
    }
}

impl Clone for FirPfbCh2 {
    fn clone(&self) -> Self {
        let mut q = std::ptr::null_mut();
        check(unsafe { sys::yg_firpfbch2_crcf_clone(self.q, &mut q) }).expect("clone failed");
        Self { q, type_: self.type_, num_channels: self.num_channels, m: self.m }
    }
}

impl Drop for FirPfbCh2 {
    fn drop(&mut self) {
        unsafe { sys::yg_firpfbch2_crcf_destroy(self.q) };
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use test_macro::autotest_annotate;

    fn reconstruction(num_channels: usize) {
        let (m, as_, tol) = (5usize, 60.0f32, 1e-3f32);
        let num_blocks = 8 * m * 2;
        let n = num_blocks * num_channels / 2;
        let mut s = 1u32;
        let x: Vec<Complex32> = (0..n)
            .map(|_| {
                s = (s * 524287) % 1031;
                Complex32::from_polar(1.0, 2.0 * std::f32::consts::PI * s as f32 / 1031.0)
            })
            .collect();
        let mut qa = FirPfbCh2::new_kaiser(FirPfbChType::Analyzer, num_channels, m, as_).unwrap();
        let mut qs = FirPfbCh2::new_kaiser(FirPfbChType::Synthesizer, num_channels, m, as_).unwrap();
        let mut ch = vec![Complex32::default(); 2 * n];
        let mut y = vec![Complex32::default(); n];
        qa.execute_block(&x, num_blocks, &mut ch).unwrap();
        qs.execute_block(&ch, num_blocks, &mut y).unwrap();
        let delay = 2 * num_channels * m - num_channels / 2 + 1;
        for i in 0..n {
            let want = if i < delay { Complex32::default() } else { x[i - delay] };
            assert!((y[i] - want).norm() < tol);
        }
    }

    #[test]
    #[autotest_annotate(autotest_firpfbch2_crcf_n8)]
    fn test_firpfbch2_crcf_n8() {
        reconstruction(8);
    }
    #[test]
    #[autotest_annotate(autotest_firpfbch2_crcf_n16)]
    fn test_firpfbch2_crcf_n16() {
        reconstruction(16);
    }
    #[test]
    #[autotest_annotate(autotest_firpfbch2_crcf_n32)]
    fn test_firpfbch2_crcf_n32() {
        reconstruction(32);
    }
    #[test]
    #[autotest_annotate(autotest_firpfbch2_crcf_n64)]
    fn test_firpfbch2_crcf_n64() {
        reconstruction(64);
    }

    #[test]
    #[autotest_annotate(autotest_firpfbch2_crcf_copy)]
    fn test_firpfbch2_crcf_copy() {
        let mut q = FirPfbCh2::new_kaiser(FirPfbChType::Analyzer, 16, 4, 60.0).unwrap();
        let x: Vec<Complex32> = (0..8).map(|i| Complex32::new(i as f32, -(i as f32))).collect();
        let mut y0 = vec![Complex32::default(); 16];
        let mut y1 = vec![Complex32::default(); 16];
        for _ in 0..7 {
            q.execute(&x, &mut y0).unwrap();
        }
        let mut c = q.clone();
        for _ in 0..24 {
            q.execute(&x, &mut y0).unwrap();
            c.execute(&x, &mut y1).unwrap();
            assert_eq!(y0, y1);
        }
    }

    #[test]
    #[autotest_annotate(autotest_firpfbch2_crcf_config)]
    fn test_firpfbch2_crcf_config() {
        assert!(FirPfbCh2::new_kaiser(FirPfbChType::Analyzer, 0, 12, 60.0).is_err());
        assert!(FirPfbCh2::new_kaiser(FirPfbChType::Analyzer, 17, 12, 60.0).is_err());
        assert!(FirPfbCh2::new_kaiser(FirPfbChType::Analyzer, 76, 0, 60.0).is_err());
        let q = FirPfbCh2::new_kaiser(FirPfbChType::Analyzer, 76, 12, 60.0).unwrap();
        assert_eq!(q.get_type(), FirPfbChType::Analyzer);
        assert_eq!(q.get_num_channels(), 76);
        assert_eq!(q.get_m(), 12);
    }
}

/// Critically sampled polyphase filterbank channelizer (upstream firpfbch_crcf), GPU-backed,
/// batched over `n_streams` independent streams that share the taps (layout x[stream][frame][M]).
#[derive(Debug)]
pub struct FirPfbCh {
    q: sys::yg_firpfbch_crcf,
    type_: FirPfbChType,
    num_channels: usize,
    p: usize,
    n_streams: usize,
}

unsafe impl Send for FirPfbCh {}

impl FirPfbCh {
    pub fn new(type_: FirPfbChType, num_channels: usize, p: usize, h: &[f32], n_streams: usize) -> Result<Self> {
        let mut q = std::ptr::null_mut();
        check(unsafe {
            sys::yg_firpfbch_crcf_create(type_ as i32, num_channels as u32, p as u32, h.as_ptr(), h.len(), n_streams as u32, &mut q)
        })?;
        Ok(Self { q, type_, num_channels, p, n_streams })
    }

    pub fn new_kaiser(type_: FirPfbChType, num_channels: usize, m: usize, as_: f32, n_streams: usize) -> Result<Self> {
        let mut q = std::ptr::null_mut();
        check(unsafe {
            sys::yg_firpfbch_crcf_create_kaiser(type_ as i32, num_channels as u32, m as u32, as_, n_streams as u32, &mut q)
        })?;
        Ok(Self { q, type_, num_channels, p: 2 * m, n_streams })
    }

    pub fn reset(&mut self) {
        let _ = unsafe { sys::yg_firpfbch_crcf_reset(self.q) };
    }
    pub fn get_type(&self) -> FirPfbChType {
        self.type_
    }
    pub fn get_num_channels(&self) -> usize {
        self.num_channels
    }
    pub fn get_p(&self) -> usize {
        self.p
    }

    /// `n` frames per stream: x and y hold n_streams * n * num_channels samples.
    pub fn execute_block(&mut self, x: &[Complex32], n: usize, y: &mut [Complex32]) -> Result<()> {
        let total = self.n_streams * n * self.num_channels;
        if x.len() != total || y.len() != total {
            return Err(Error::Config("input/output block lengths do not match the frame count".into()));
        }
        check(unsafe {
            sys::yg_firpfbch_crcf_execute_block(self.q, x.as_ptr() as *const sys::yg_cf32, n, y.as_mut_ptr() as *mut sys::yg_cf32)
        })
    }
}

impl Clone for FirPfbCh {
    fn clone(&self) -> Self {
        let mut q = std::ptr::null_mut();
        check(unsafe { sys::yg_firpfbch_crcf_clone(self.q, &mut q) }).expect("clone failed");
        Self { q, type_: self.type_, num_channels: self.num_channels, p: self.p, n_streams: self.n_streams }
    }
}

impl Drop for FirPfbCh {
    fn drop(&mut self) {
        unsafe { sys::yg_firpfbch_crcf_destroy(self.q) };
    }
}
