//! heimdall-cuda -- safe Rust wrapper over the C ABI of `include/heimdall_cuda.h`.
//!
//! SOURCE ONLY: not compiled or tested in this environment (no cargo/rustc).  The `extern "C"` block is a one-to-one
//! transcription of the header; the wrapper gives the signature of `heimdall_core::detection::detect_contamination`
//! (rust/heimdall-core/src/detection.rs:127-132) so that `lib.rs:111-113` can switch to it under a `cuda` feature.
#![allow(non_camel_case_types)]

use ndarray::ArrayView3;
use std::ffi::CStr;
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct hv_ctx {
    _private: [u8; 0],
}
pub type hv_status = i32;
pub const HV_OK: hv_status = 0;
pub const HV_ERR_INVALID_DIMENSIONS: hv_status = -1;
pub const HV_ERR_CAPACITY: hv_status = -4;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct hv_config {
    pub max_batch: i32,
    pub max_height: i32,
    pub max_width: i32,
    pub max_blobs_per_frame: i32,
    pub max_defects_per_frame: i32,
    pub num_slots: i32,
    pub flags: i32,
    pub reserved: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct hv_params {
    pub min_size: f64,
    pub max_size: f64,
    pub threshold: f64,
    pub min_confidence: f64,
    pub gauss_sigma: f64,
    pub blur_mode: i32,
    pub blur_ksize: i32,
    pub morph_open_k: i32,
    pub morph_close_k: i32,
    pub reserved: [i32; 4],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct hv_defect {
    pub y: i32,
    pub x: i32,
    pub size: f64,
    pub confidence: f64,
    pub ymin: i32,
    pub xmin: i32,
    pub ymax: i32,
    pub xmax: i32,
    pub label: u32,
    pub frame: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct hv_frame_result {
    pub n_components: u32,
    pub n_defects: u32,
    pub defects_offset: u32,
    pub rejected: u32,
    pub fg_pixels: u32,
    pub status: i32,
}

extern "C" {
    pub fn hv_params_default(p: *mut hv_params);
    pub fn hv_create(device: i32, cfg: *const hv_config, out: *mut *mut hv_ctx) -> hv_status;
    pub fn hv_destroy(ctx: *mut hv_ctx);
    pub fn hv_last_error(ctx: *const hv_ctx) -> *const c_char;
    pub fn hv_host_alloc(ctx: *mut hv_ctx, bytes: usize) -> *mut c_void;
    pub fn hv_host_free(ctx: *mut hv_ctx, p: *mut c_void);
    /// Device buffers for the device-resident entry points; flags = HV_ALLOC_COMPRESSIBLE (1) asks for L2
    /// compute-data compression (the mostly-zero mask / label planes then cost less DRAM write time).
    pub fn hv_pipeline_depth() -> i32;
    pub fn hv_device_alloc(ctx: *mut hv_ctx, bytes: usize, flags: u32, d_ptr: *mut *mut c_void, compressed_out: *mut i32) -> hv_status;
    pub fn hv_device_free(ctx: *mut hv_ctx, d_ptr: *mut c_void) -> hv_status;
    pub fn hv_device_read(ctx: *mut hv_ctx, host_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> hv_status;
    pub fn hv_device_write(ctx: *mut hv_ctx, d_dst: *mut c_void, host_src: *const c_void, bytes: usize) -> hv_status;
    pub fn hv_detect_batch(
        ctx: *mut hv_ctx,
        frames: *const u8,
        n: i32,
        h: i32,
        w: i32,
        c: i32,
        row_stride: usize,
        frame_stride: usize,
        params: *const hv_params,
        results: *mut hv_frame_result,
        defects: *mut hv_defect,
        defects_cap: usize,
        n_defects_total: *mut usize,
        debug: *const c_void,
    ) -> hv_status;
    pub fn hv_submit(
        ctx: *mut hv_ctx,
        frames: *const u8,
        n: i32,
        h: i32,
        w: i32,
        c: i32,
        row_stride: usize,
        frame_stride: usize,
        params: *const hv_params,
        ticket: *mut i64,
    ) -> hv_status;
    pub fn hv_wait(
        ctx: *mut hv_ctx,
        ticket: i64,
        results: *mut hv_frame_result,
        defects: *mut hv_defect,
        defects_cap: usize,
        n_defects_total: *mut usize,
    ) -> hv_status;
    pub fn hv_stats_device_ptr(ctx: *mut hv_ctx) -> *mut u64;
}

/// Same fields as `heimdall_core::detection::Defect` (detection.rs:12-18); metadata is filled by the Python layer.
#[derive(Debug, Clone)]
pub struct Defect {
    pub position: (usize, usize),
    pub size: f64,
    pub confidence: f64,
}

#[derive(thiserror::Error, Debug)]
pub enum DetectionError {
    #[error("Detection error: {0}")]
    Detection(String),
    #[error("Invalid image dimensions: expected 3D array")]
    InvalidDimensions,
}

/// One CUDA device, one context.  Not `Sync`: use one per thread, as the C ABI requires.
pub struct CudaDetector {
    ctx: *mut hv_ctx,
}

impl CudaDetector {
    pub fn new(device: i32) -> Result<Self, DetectionError> {
        let mut ctx = std::ptr::null_mut();
        let cfg = hv_config::default();
        let st = unsafe { hv_create(device, &cfg, &mut ctx) };
        if st != HV_OK {
            let msg = unsafe { CStr::from_ptr(hv_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(DetectionError::Detection(msg));
        }
        Ok(Self { ctx })
    }

    /// Drop-in for `detection::detect_contamination(image, min_size, max_size, threshold)`.
    pub fn detect(&self, image: &ArrayView3<u8>, min_size: f64, max_size: f64, threshold: f64) -> Result<Vec<Defect>, DetectionError> {
        let (h, w, c) = image.dim();
        if c != 1 && c != 3 {
            return Err(DetectionError::InvalidDimensions);
        }
        let owned;
        let data: &[u8] = match image.as_slice() {
            Some(s) => s,
            None => {
                owned = image.to_owned();
                owned.as_slice().expect("contiguous after to_owned")
            }
        };
        let mut p = std::mem::MaybeUninit::<hv_params>::uninit();
        let mut p = unsafe {
            hv_params_default(p.as_mut_ptr());
            p.assume_init()
        };
        p.min_size = min_size;
        p.max_size = max_size;
        p.threshold = threshold;
        let mut res = hv_frame_result::default();
        let cap = 256usize;
        let mut defects: Vec<hv_defect> = Vec::with_capacity(cap);
        let mut n = 0usize;
        let st = unsafe {
            hv_detect_batch(self.ctx, data.as_ptr(), 1, h as i32, w as i32, c as i32, 0, 0, &p, &mut res, defects.as_mut_ptr(), cap, &mut n, std::ptr::null())
        };
        if st == HV_ERR_INVALID_DIMENSIONS {
            return Err(DetectionError::InvalidDimensions);
        }
        if st != HV_OK {
            let msg = unsafe { CStr::from_ptr(hv_last_error(self.ctx)) }.to_string_lossy().into_owned();
            return Err(DetectionError::Detection(msg));
        }
        unsafe { defects.set_len(n) };
        Ok(defects
            .iter()
            .map(|d| Defect { position: (d.y as usize, d.x as usize), size: d.size, confidence: d.confidence })
            .collect())
    }
}

impl Drop for CudaDetector {
    fn drop(&mut self) {
        unsafe { hv_destroy(self.ctx) }
    }
}
