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

/// A device buffer owned by a context (`zkb_dev_alloc` / `zkb_dev_free`).
pub struct DevBuf<'a> {
    ctx: &'a Ctx,
    ptr: *mut core::ffi::c_void,
}

impl DevBuf<'_> {
    pub fn ptr(&self) -> *const u64 {
        self.ptr as *const u64
    }
    pub fn ptr_mut(&self) -> *mut u64 {
        self.ptr as *mut u64
    }
}

impl Drop for DevBuf<'_> {
    fn drop(&mut self) {
        unsafe { sys::zkb_dev_free(self.ctx.raw, self.ptr) };
    }
}

impl Ctx {
    /// `bytes` of host memory copied into a fresh device buffer (synchronous).
    pub fn upload(&self, src: *const u64, bytes: usize) -> Result<DevBuf<'_>, Error> {
        let mut ptr = core::ptr::null_mut();
        self.check(unsafe { sys::zkb_dev_alloc(self.raw, bytes, &mut ptr) })?;
        let buf = DevBuf { ctx: self, ptr };
        self.check(unsafe { sys::zkb_h2d(self.raw, ptr, src as *const core::ffi::c_void, bytes) })?;
        Ok(buf)
    }

    pub fn download(&self, dst: *mut u64, src_dev: *const u64, bytes: usize) -> Result<(), Error> {
        self.check(unsafe { sys::zkb_d2h(self.raw, dst as *mut core::ffi::c_void, src_dev as *const core::ffi::c_void, bytes) })
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
