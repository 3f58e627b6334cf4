//! GPU-backed `FirFilter<Complex32, f32>` (crcf) over libyagi_b200, batched over `n_streams`
//! independent streams that share the taps.  Mirrors the method names and error behaviour of
//! src/filter/fir/firfilt.rs (`new` :63-79, `new_kaiser` :93-110, `reset` :209-211,
//! `execute_block` :267-278, `set_scale`/`get_scale` :285-296, `get_length` :303-305); one
//! object with `n_streams = S` stands for S reference objects run side by side.
//!
//! SOURCE ONLY: not compiled in the development image (no Rust toolchain there).
use crate::error::{Error, Result};
use num_complex::Complex32;
use std::ffi::CStr;
use yagi_b200_sys as sys;

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

#[derive(Debug)]
pub struct FirFiltGpu {
    q: sys::yg_firfilt_crcf,
    h_len: usize,
    n_streams: usize,
}

unsafe impl Send for FirFiltGpu {}

impl FirFiltGpu {
    /// `FirFilter::new(h)`: `Err(Error::Config)` for an empty filter (firfilt.rs:64-66).
    pub fn new(h: &[f32], n_streams: usize) -> Result<Self> {
        let mut q = std::ptr::null_mut();
        check(unsafe { sys::yg_firfilt_crcf_create(h.as_ptr(), h.len(), n_streams as u32, &mut q) })?;
        Ok(Self { q, h_len: h.len(), n_streams })
    }

    /// `FirFilter::new_kaiser(n, fc, as_, mu)` (firfilt.rs:93-110).
    pub fn new_kaiser(n: usize, fc: f32, as_: f32, mu: f32, n_streams: usize) -> Result<Self> {
        let mut q = std::ptr::null_mut();
        check(unsafe { sys::yg_firfilt_crcf_create_kaiser(n as u32, fc, as_, mu, n_streams as u32, &mut q) })?;
        Ok(Self { q, h_len: n, n_streams })
    }

    pub fn reset(&mut self) {
        let _ = unsafe { sys::yg_firfilt_crcf_reset(self.q) };
    }

    pub fn set_scale(&mut self, scale: f32) {
        let _ = unsafe { sys::yg_firfilt_crcf_set_scale(self.q, scale) };
    }

    pub fn get_scale(&self) -> f32 {
        let mut s = 0.0f32;
        let _ = unsafe { sys::yg_firfilt_crcf_get_scale(self.q, &mut s) };
        s
    }

    pub fn get_length(&self) -> usize {
        self.h_len
    }

    pub fn get_num_streams(&self) -> usize {
        self.n_streams
    }

    /// x[stream][n] -> y[stream][n]; `Err(Error::Range)`-style length check as firfilt.rs:268-270.
    pub fn execute_block(&mut self, x: &[Complex32], y: &mut [Complex32]) -> Result<()> {
        if x.len() != y.len() || self.n_streams == 0 || x.len() % self.n_streams != 0 {
            return Err(Error::Range("input and output blocks must have the same length, a multiple of n_streams".into()));
        }
        let n = x.len() / self.n_streams;
        check(unsafe {
            sys::yg_firfilt_crcf_execute_block(self.q, x.as_ptr() as *const sys::yg_cf32, n, y.as_mut_ptr() as *mut sys::yg_cf32)
        })
    }
}

impl Clone for FirFiltGpu {
    fn clone(&self) -> Self {
        let mut q = std::ptr::null_mut();
        check(unsafe { sys::yg_firfilt_crcf_clone(self.q, &mut q) }).expect("clone failed");
        Self { q, h_len: self.h_len, n_streams: self.n_streams }
    }
}

impl Drop for FirFiltGpu {
    fn drop(&mut self) {
        unsafe { sys::yg_firfilt_crcf_destroy(self.q) };
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use test_macro::autotest_annotate;

    // the reference's own vectors (src/filter/fir/firfilt_test_data.rs) through the GPU object
    use crate::filter::fir::firfilt_test_data::*;

    fn run(h: &[f32], x: &[Complex32], want: &[Complex32]) {
        let mut q = FirFiltGpu::new(h, 1).unwrap();
        let mut y = vec![Complex32::default(); x.len()];
        q.execute_block(x, &mut y).unwrap();
        for (a, b) in y.iter().zip(want) {
            assert!((a - b).norm() < 1e-3);
        }
    }

    #[test]
    #[autotest_annotate(autotest_firfilt_crcf_data_h4x8)]
    fn test_firfilt_crcf_gpu_h4x8() {
        run(&FIRFILT_CRCF_DATA_H4X8_H, &FIRFILT_CRCF_DATA_H4X8_X, &FIRFILT_CRCF_DATA_H4X8_Y);
    }

    #[test]
    fn test_firfilt_crcf_gpu_config() {
        assert!(FirFiltGpu::new(&[], 1).is_err());
        assert!(FirFiltGpu::new(&[1.0], 0).is_err());
    }
}
