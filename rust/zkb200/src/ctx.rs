//! One `zkb_ctx` per proving thread and GPU (`prove()` is `!Send`: plonk-core/src/proof_system/prove.rs:62).
use core::ffi::c_int;
use std::ffi::CStr;
use zkb200_sys as sys;

#[derive(Debug)]
pub struct Error {
    pub code: c_int,
    pub message: String,
}

pub struct Ctx {
    raw: *mut sys::zkb_ctx,
}

impl Ctx {
    /// `zkb_ctx_create`: fails when no CUDA device is present -- there is no CPU fallback.
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut raw = core::ptr::null_mut();
        let rc = unsafe { sys::zkb_ctx_create(device, &mut raw) };
        if rc != sys::ZKB_OK {
            return Err(Error { code: rc, message: "zkb_ctx_create failed (no CUDA device?)".into() });
        }
        Ok(Self { raw })
    }

    pub fn raw(&self) -> *mut sys::zkb_ctx {
        self.raw
    }

    pub fn check(&self, rc: c_int) -> Result<(), Error> {
        if rc == sys::ZKB_OK {
            return Ok(());
        }
        let message = unsafe { CStr::from_ptr(sys::zkb_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message })
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::zkb_ctx_destroy(self.raw) }
    }
}

thread_local! {
    /// The proving thread's context on GPU `ZKB200_DEVICE` (default 0).
    pub static CTX: Ctx = Ctx::new(std::env::var("ZKB200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0))
        .expect("zkb200: cannot create a CUDA context");
}
