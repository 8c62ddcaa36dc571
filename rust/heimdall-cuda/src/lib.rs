//! heimdall-cuda -- safe Rust wrapper over the C ABI of `include/heimdall_cuda.h`.
//!
//! SOURCE ONLY: not compiled or tested in this environment (no cargo/rustc).  `ffi.rs` (constants, `#[repr(C)]` structs and
//! the complete `extern "C"` block, one declaration per `HV_API` symbol) is GENERATED from the header by
//! `tools/gen_rust_ffi.py`, and `tests/test_abi.py` fails when it is out of date.  The wrapper below gives the signature of
//! `heimdall_core::detection::detect_contamination` (rust/heimdall-core/src/detection.rs:127-132) so that
//! `rust/heimdall-core/src/lib.rs:111-113` can switch to it under a `cuda` feature, plus the batched streaming form a
//! camera loop uses (heimdall-camera / heimdall-gige frames -> pinned staging -> `hv_submit` / `hv_wait`).
//!
//! Multi-GPU: the C ABI is per device.  A host that shards camera streams over several GPUs creates one `CudaDetector`
//! per device (stream s -> device s % G, rust/heimdall-gige/src/lib.rs:584-616 is the per-camera fan-out it replaces) and
//! owns the NCCL communicator itself: `stats_device_ptr()` is the 32 x u64 vector to pass to
//! `ncclAllReduce(ncclUint64, ncclSum)` on a side stream.  This crate ships no sharding helper.
#![allow(non_camel_case_types)]

pub mod ffi;
pub use ffi::*;

use ndarray::ArrayView3;
use std::ffi::CStr;

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
    defect_cap: usize,
}

fn default_params(min_size: f64, max_size: f64, threshold: f64) -> hv_params {
    let mut p = std::mem::MaybeUninit::<hv_params>::uninit();
    let mut p = unsafe {
        hv_params_default(p.as_mut_ptr());
        p.assume_init()
    };
    p.min_size = min_size;
    p.max_size = max_size;
    p.threshold = threshold;
    p
}

impl CudaDetector {
    pub fn new(device: i32) -> Result<Self, DetectionError> {
        let mut ctx = std::ptr::null_mut();
        let mut cfg = std::mem::MaybeUninit::<hv_config>::uninit();
        let cfg = unsafe {
            hv_config_default(cfg.as_mut_ptr());
            cfg.assume_init()
        };
        let st = unsafe { hv_create(device, &cfg, &mut ctx) };
        if st != HV_OK {
            let msg = unsafe { CStr::from_ptr(hv_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(DetectionError::Detection(msg));
        }
        Ok(Self { ctx, defect_cap: 256 })
    }

    fn error(&self, st: hv_status) -> DetectionError {
        if st == HV_ERR_INVALID_DIMENSIONS {
            return DetectionError::InvalidDimensions;
        }
        let msg = unsafe { CStr::from_ptr(hv_last_error(self.ctx)) }.to_string_lossy().into_owned();
        DetectionError::Detection(msg)
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
        let p = default_params(min_size, max_size, threshold);
        let mut res = hv_frame_result { n_components: 0, n_defects: 0, defects_offset: 0, rejected: 0, fg_pixels: 0, status: 0 };
        let cap = self.defect_cap;
        let mut defects: Vec<hv_defect> = Vec::with_capacity(cap);
        let mut n = 0usize;
        let st = unsafe {
            hv_detect_batch(self.ctx, data.as_ptr(), 1, h as i32, w as i32, c as i32, 0, 0, &p, &mut res, defects.as_mut_ptr(), cap, &mut n, std::ptr::null())
        };
        if st != HV_OK {
            return Err(self.error(st));
        }
        unsafe { defects.set_len(n) };
        Ok(defects
            .iter()
            .map(|d| Defect { position: (d.y as usize, d.x as usize), size: d.size, confidence: d.confidence })
            .collect())
    }

    /// Asynchronous batch from page-locked host memory (`hv_host_alloc`): returns a ticket for `wait`.
    /// `frames` holds `n` tightly packed `h x w` Mono8 frames and must stay valid until `wait` returns.
    pub fn submit(&self, frames: &[u8], n: usize, h: usize, w: usize, min_size: f64, max_size: f64, threshold: f64) -> Result<i64, DetectionError> {
        assert!(frames.len() >= n * h * w);
        let p = default_params(min_size, max_size, threshold);
        let mut ticket = 0i64;
        let st = unsafe { hv_submit(self.ctx, frames.as_ptr(), n as i32, h as i32, w as i32, 1, 0, 0, &p, &mut ticket) };
        if st != HV_OK {
            return Err(self.error(st));
        }
        Ok(ticket)
    }

    /// Per-frame reject decisions (`has_defects`, heimdall/inspection/base_inspector.py:40-42) and defect lists of a batch.
    pub fn wait(&self, ticket: i64, n: usize) -> Result<Vec<(bool, Vec<Defect>)>, DetectionError> {
        let mut res = vec![hv_frame_result { n_components: 0, n_defects: 0, defects_offset: 0, rejected: 0, fg_pixels: 0, status: 0 }; n];
        let cap = n * self.defect_cap;
        let mut defects: Vec<hv_defect> = Vec::with_capacity(cap);
        let mut total = 0usize;
        let st = unsafe { hv_wait(self.ctx, ticket, res.as_mut_ptr(), defects.as_mut_ptr(), cap, &mut total) };
        if st != HV_OK {
            return Err(self.error(st));
        }
        unsafe { defects.set_len(total) };
        Ok(res
            .iter()
            .map(|r| {
                let o = r.defects_offset as usize;
                let list = defects[o..o + r.n_defects as usize]
                    .iter()
                    .map(|d| Defect { position: (d.y as usize, d.x as usize), size: d.size, confidence: d.confidence })
                    .collect();
                (r.rejected != 0, list)
            })
            .collect())
    }

    /// Device pointer of the 32 x u64 line statistics, for the host's own `ncclAllReduce`.
    pub fn stats_device_ptr(&self) -> *mut u64 {
        unsafe { hv_stats_device_ptr(self.ctx) }
    }
}

impl Drop for CudaDetector {
    fn drop(&mut self) {
        unsafe { hv_destroy(self.ctx) }
    }
}
