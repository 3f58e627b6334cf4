//! Raw bindings to include/yagi_b200.h (one `extern "C"` item per declared symbol).
#![allow(non_camel_case_types)]
use libc::{c_char, c_void, size_t};

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct yg_cf32 {
    pub re: f32,
    pub im: f32,
}

#[repr(C)]
pub struct yg_firpfbch2_crcf_s {
    _private: [u8; 0],
}
#[repr(C)]
pub struct yg_firpfbch_crcf_s {
    _private: [u8; 0],
}
#[repr(C)]
pub struct yg_firfilt_crcf_s {
    _private: [u8; 0],
}
pub type yg_firpfbch2_crcf = *mut yg_firpfbch2_crcf_s;
pub type yg_firpfbch_crcf = *mut yg_firpfbch_crcf_s;
pub type yg_firfilt_crcf = *mut yg_firfilt_crcf_s;

pub const YG_OK: i32 = 0;
pub const YG_EINTERNAL: i32 = 1;
pub const YG_ECONFIG: i32 = 2;
pub const YG_EVALUE: i32 = 3;
pub const YG_ERANGE: i32 = 4;
pub const YG_EMODE: i32 = 5;
pub const YG_ENOCONV: i32 = 6;
pub const YG_ANALYZER: i32 = 0;
pub const YG_SYNTHESIZER: i32 = 1;

extern "C" {
    pub fn yg_version() -> i32;
    pub fn yg_last_error() -> *const c_char;
    pub fn yg_device_count(n: *mut i32) -> i32;
    pub fn yg_launch_count(n: *mut u64) -> i32;
    pub fn yg_channel_major_dev(d_frames: *const yg_cf32, n_frames: size_t, m_ch: u32, d_out: *mut yg_cf32, cuda_stream: *mut c_void) -> i32;
    pub fn yg_host_alloc(p: *mut *mut c_void, bytes: size_t) -> i32;
    pub fn yg_host_free(p: *mut c_void) -> i32;
    pub fn yg_fir_design_kaiser(n: u32, fc: f32, as_: f32, mu: f32, h: *mut f32) -> i32;

    pub fn yg_firpfbch2_crcf_create(type_: i32, m_ch: u32, m: u32, h: *const f32, h_len: size_t, out: *mut yg_firpfbch2_crcf) -> i32;
    pub fn yg_firpfbch2_crcf_create_kaiser(type_: i32, m_ch: u32, m: u32, as_: f32, out: *mut yg_firpfbch2_crcf) -> i32;
    pub fn yg_firpfbch2_crcf_clone(q: yg_firpfbch2_crcf, out: *mut yg_firpfbch2_crcf) -> i32;
    pub fn yg_firpfbch2_crcf_destroy(q: yg_firpfbch2_crcf) -> i32;
    pub fn yg_firpfbch2_crcf_reset(q: yg_firpfbch2_crcf) -> i32;
    pub fn yg_firpfbch2_crcf_execute(q: yg_firpfbch2_crcf, x: *const yg_cf32, y: *mut yg_cf32) -> i32;
    pub fn yg_firpfbch2_crcf_execute_block(q: yg_firpfbch2_crcf, x: *const yg_cf32, n_frames: size_t, y: *mut yg_cf32) -> i32;
    pub fn yg_firpfbch2_crcf_execute_block_dev(q: yg_firpfbch2_crcf, d_x: *const yg_cf32, n_frames: size_t, d_y: *mut yg_cf32, cuda_stream: *mut c_void) -> i32;
    pub fn yg_firpfbch2_crcf_sync(q: yg_firpfbch2_crcf) -> i32;
    pub fn yg_firpfbch2_crcf_get_type(q: yg_firpfbch2_crcf, type_: *mut i32) -> i32;
    pub fn yg_firpfbch2_crcf_get_M(q: yg_firpfbch2_crcf, m_ch: *mut u32) -> i32;
    pub fn yg_firpfbch2_crcf_get_m(q: yg_firpfbch2_crcf, m: *mut u32) -> i32;
    pub fn yg_firpfbch2_crcf_get_taps(q: yg_firpfbch2_crcf, h: *mut f32) -> i32;
    pub fn yg_firpfbch2_crcf_get_device(q: yg_firpfbch2_crcf, dev: *mut i32) -> i32;
    pub fn yg_firpfbch2_crcf_state_len(q: yg_firpfbch2_crcf, n: *mut size_t) -> i32;
    pub fn yg_firpfbch2_crcf_get_state(q: yg_firpfbch2_crcf, hist: *mut yg_cf32, flag: *mut i32) -> i32;
    pub fn yg_firpfbch2_crcf_set_state(q: yg_firpfbch2_crcf, hist: *const yg_cf32, flag: i32) -> i32;
    pub fn yg_firpfbch2_crcf_last_path(q: yg_firpfbch2_crcf, path: *mut i32) -> i32;
    pub fn yg_firpfbch2_crcf_last_kernel_ms(q: yg_firpfbch2_crcf, ms: *mut f32) -> i32;
    pub fn yg_firpfbch2_crcf_kernel_times(q: yg_firpfbch2_crcf, ms: *mut f32, cap: size_t, n: *mut size_t) -> i32;
    pub fn yg_firpfbch2_crcf_set_kernel_timing(q: yg_firpfbch2_crcf, enable: i32) -> i32;

    pub fn yg_firpfbch_crcf_create(type_: i32, m_ch: u32, p: u32, h: *const f32, h_len: size_t, n_streams: u32, out: *mut yg_firpfbch_crcf) -> i32;
    pub fn yg_firpfbch_crcf_create_kaiser(type_: i32, m_ch: u32, m: u32, as_: f32, n_streams: u32, out: *mut yg_firpfbch_crcf) -> i32;
    pub fn yg_firpfbch_crcf_clone(q: yg_firpfbch_crcf, out: *mut yg_firpfbch_crcf) -> i32;
    pub fn yg_firpfbch_crcf_destroy(q: yg_firpfbch_crcf) -> i32;
    pub fn yg_firpfbch_crcf_reset(q: yg_firpfbch_crcf) -> i32;
    pub fn yg_firpfbch_crcf_execute(q: yg_firpfbch_crcf, x: *const yg_cf32, y: *mut yg_cf32) -> i32;
    pub fn yg_firpfbch_crcf_execute_block(q: yg_firpfbch_crcf, x: *const yg_cf32, n_frames: size_t, y: *mut yg_cf32) -> i32;
    pub fn yg_firpfbch_crcf_execute_block_dev(q: yg_firpfbch_crcf, d_x: *const yg_cf32, n_frames: size_t, d_y: *mut yg_cf32, cuda_stream: *mut c_void) -> i32;
    pub fn yg_firpfbch_crcf_sync(q: yg_firpfbch_crcf) -> i32;
    pub fn yg_firpfbch_crcf_get_type(q: yg_firpfbch_crcf, type_: *mut i32) -> i32;
    pub fn yg_firpfbch_crcf_get_M(q: yg_firpfbch_crcf, m_ch: *mut u32) -> i32;
    pub fn yg_firpfbch_crcf_get_p(q: yg_firpfbch_crcf, p: *mut u32) -> i32;
    pub fn yg_firpfbch_crcf_get_n_streams(q: yg_firpfbch_crcf, n: *mut u32) -> i32;
    pub fn yg_firpfbch_crcf_get_taps(q: yg_firpfbch_crcf, h: *mut f32) -> i32;
    pub fn yg_firpfbch_crcf_get_device(q: yg_firpfbch_crcf, dev: *mut i32) -> i32;
    pub fn yg_firpfbch_crcf_last_path(q: yg_firpfbch_crcf, path: *mut i32) -> i32;

    pub fn yg_firfilt_crcf_create(h: *const f32, h_len: size_t, n_streams: u32, out: *mut yg_firfilt_crcf) -> i32;
    pub fn yg_firfilt_crcf_create_kaiser(n: u32, fc: f32, as_: f32, mu: f32, n_streams: u32, out: *mut yg_firfilt_crcf) -> i32;
    pub fn yg_firfilt_crcf_clone(q: yg_firfilt_crcf, out: *mut yg_firfilt_crcf) -> i32;
    pub fn yg_firfilt_crcf_destroy(q: yg_firfilt_crcf) -> i32;
    pub fn yg_firfilt_crcf_reset(q: yg_firfilt_crcf) -> i32;
    pub fn yg_firfilt_crcf_set_scale(q: yg_firfilt_crcf, scale: f32) -> i32;
    pub fn yg_firfilt_crcf_get_scale(q: yg_firfilt_crcf, scale: *mut f32) -> i32;
    pub fn yg_firfilt_crcf_get_len(q: yg_firfilt_crcf, h_len: *mut size_t) -> i32;
    pub fn yg_firfilt_crcf_get_device(q: yg_firfilt_crcf, dev: *mut i32) -> i32;
    pub fn yg_firfilt_crcf_last_path(q: yg_firfilt_crcf, path: *mut i32) -> i32;
    pub fn yg_firfilt_crcf_execute_block(q: yg_firfilt_crcf, x: *const yg_cf32, n: size_t, y: *mut yg_cf32) -> i32;
    pub fn yg_firfilt_crcf_execute_block_dev(q: yg_firfilt_crcf, d_x: *const yg_cf32, n: size_t, d_y: *mut yg_cf32, cuda_stream: *mut c_void) -> i32;
    pub fn yg_firfilt_crcf_sync(q: yg_firfilt_crcf) -> i32;
}
