//! `rs_tfhe::cuda` -- thin FFI over include/tfhe_b200.h.
//!
//! * `CudaBootstrap` implements `bootstrap::Bootstrap` (src/bootstrap/mod.rs:23-38), the slot
//!   examples/bootstrap_strategies.rs:94-97 reserves for `"gpu"`.
//! * `batch_*` mirror gates::batch_*[_with_railgun] (src/gates.rs:352-547) and
//!   trgsw::batch_blind_rotate (src/trgsw.rs:289-305).
//! * `LutBootstrap::bootstrap_lut` / `bootstrap_func` (src/bootstrap/lut.rs:49-99) get a batched
//!   device path.
//!
//! Lives inside the crate because `TRGSWLv1FFT.trlwe_fft` is private (src/trgsw.rs:53-55); the
//! bootstrapping key is passed as the plain nested array it is:
//! `[TRGSWLv1FFT; n]` == `f64[n][2l][2][1024]`.
use std::os::raw::{c_char, c_double, c_int, c_void};

use crate::bootstrap::Bootstrap;
use crate::key::CloudKey;
use crate::params;
use crate::trlwe::TRLWELv1;
use crate::utils::Ciphertext;

#[repr(C)]
struct TfheParams { n: u32, big_n: u32, l: u32, bgbit: u32, basebit: u32, iks_t: u32 }
#[repr(C)]
pub struct TfheEngine { _private: [u8; 0] }

#[link(name = "tfhe_b200")]
extern "C" {
    fn tfhe_last_error() -> *const c_char;
    fn tfhe_engine_create(p: *const TfheParams, device: c_int, out: *mut *mut TfheEngine) -> c_int;
    fn tfhe_engine_create_multi(p: *const TfheParams, device_ids: *const c_int, n_devices: c_int,
        out: *mut *mut TfheEngine) -> c_int;
    fn tfhe_engine_destroy(e: *mut TfheEngine);
    fn tfhe_engine_load_cloud_key(e: *mut TfheEngine, decomposition_offset: u32,
        testvec_a: *const u32, testvec_b: *const u32, ksk: *const u32, bsk: *const c_double) -> c_int;
    fn tfhe_batch_gate(e: *mut TfheEngine, op: c_int, in_pairs: *const u32, out: *mut u32, count: usize) -> c_int;
    fn tfhe_batch_bootstrap(e: *mut TfheEngine, input: *const u32, out: *mut u32, count: usize, key_switch: c_int) -> c_int;
    fn tfhe_batch_blind_rotate(e: *mut TfheEngine, input: *const u32, out_trlwe: *mut u32, count: usize) -> c_int;
    fn tfhe_lut_generate(e: *mut TfheEngine, f_table: *const u32, modulus: u32, scale: c_double,
        lut_b_out: *mut u32, lut_id_out: *mut c_int) -> c_int;
    fn tfhe_batch_bootstrap_lut(e: *mut TfheEngine, lut_id: c_int, input: *const u32, out: *mut u32, count: usize) -> c_int;
    fn tfhe_lut_release(e: *mut TfheEngine, lut_id: c_int) -> c_int;
    fn tfhe_batch_bootstrap_func(e: *mut TfheEngine, f_table: *const u32, modulus: u32, scale: c_double,
        input: *const u32, out: *mut u32, count: usize) -> c_int;
    // FFTProcessor seam (src/fft/mod.rs:80-107)
    fn tfhe_batch_ifft(e: *mut TfheEngine, input: *const u32, out: *mut c_double, count: usize) -> c_int;
    fn tfhe_batch_fft(e: *mut TfheEngine, input: *const c_double, out: *mut u32, count: usize) -> c_int;
    fn tfhe_batch_poly_mul(e: *mut TfheEngine, a: *const u32, b: *const u32, out: *mut u32, count: usize) -> c_int;
    fn tfhe_batch_bootstrap_lut_multi(e: *mut TfheEngine, lut_ids: *const i32, input: *const u32, out: *mut u32, count: usize) -> c_int;
    // key.rs:59-66 on the device; key blob checkpointing; proxy_reenc.rs:468-511
    fn tfhe_engine_generate_cloud_key(e: *mut TfheEngine, s0: *const u32, s1: *const u32,
        alpha_lv0: c_double, alpha_lv1: c_double, seed: u64) -> c_int;
    fn tfhe_engine_cloud_key_export_bytes(e: *mut TfheEngine) -> usize;
    fn tfhe_engine_export_cloud_key(e: *mut TfheEngine, buf: *mut c_void, bytes: usize) -> c_int;
    fn tfhe_engine_import_cloud_key(e: *mut TfheEngine, buf: *const c_void, bytes: usize) -> c_int;
    fn tfhe_reenc_key_load(e: *mut TfheEngine, key_encryptions: *const u32, base: u32, t: u32,
        out: *mut *mut c_void) -> c_int;
    fn tfhe_reenc_key_destroy(k: *mut c_void);
    fn tfhe_batch_reencrypt(e: *mut TfheEngine, key: *const c_void, input: *const u32, out: *mut u32, count: usize) -> c_int;
}

#[derive(Copy, Clone)]
#[repr(i32)]
pub enum Gate { Nand = 0, And = 1, Or = 2, Xor = 3, Xnor = 4, Nor = 5, AndNy = 6, AndYn = 7, OrNy = 8, OrYn = 9 }

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(tfhe_last_error()) }.to_string_lossy().into_owned();
        panic!("tfhe_b200: {}", msg); // the reference's error model is panic (unwrap everywhere)
    }
}

/// Identity of a CloudKey's CONTENT: FNV-1a over the offset, the lengths and 1024 strided samples of
/// each array.  An address is not an identity -- a dropped key and a new one can share it.
fn fingerprint(ck: &CloudKey) -> u64 {
    let mut h: u64 = 1469598103934665603;
    let mut mix = |v: u64| for i in 0..8 { h ^= (v >> (8 * i)) & 0xff; h = h.wrapping_mul(1099511628211); };
    mix(ck.decomposition_offset as u64);
    mix(ck.key_switching_key.len() as u64);
    mix(ck.bootstrapping_key.len() as u64);
    for i in (0..params::trgsw_lv1::N).step_by(8) {
        mix(ck.blind_rotate_testvec.a[i] as u64);
        mix(ck.blind_rotate_testvec.b[i] as u64);
    }
    let ksk = unsafe { std::slice::from_raw_parts(ck.key_switching_key.as_ptr() as *const u32,
        ck.key_switching_key.len() * (params::tlwe_lv0::N + 1)) };
    let bsk = unsafe { std::slice::from_raw_parts(ck.bootstrapping_key.as_ptr() as *const u64,
        ck.bootstrapping_key.len() * 2 * params::trgsw_lv1::L * 2 * params::trgsw_lv1::N) };
    for i in 0..1024usize {
        if !ksk.is_empty() { mix(ksk[i.wrapping_mul(2654435761) % ksk.len()] as u64); }
        if !bsk.is_empty() { mix(bsk[i.wrapping_mul(2654435761) % bsk.len()]); }
    }
    h | 1   // never 0 ("no key") or the device-generated marker
}

/// One engine per (process, GPU) -- or per GPU set (`with_devices`); the cloud key stays resident on
/// the device(s).
pub struct CudaBootstrap { raw: *mut TfheEngine, loaded_key: std::sync::Mutex<u64> }
unsafe impl Send for CudaBootstrap {}
unsafe impl Sync for CudaBootstrap {} // calls are serialised inside the engine

impl CudaBootstrap {
    pub fn new(device: i32) -> Self {
        let p = TfheParams {
            n: params::tlwe_lv0::N as u32, big_n: params::trgsw_lv1::N as u32,
            l: params::trgsw_lv1::L as u32, bgbit: params::trgsw_lv1::BGBIT,
            basebit: params::trgsw_lv1::BASEBIT as u32, iks_t: params::trgsw_lv1::IKS_T as u32,
        };
        let mut raw = std::ptr::null_mut();
        check(unsafe { tfhe_engine_create(&p, device, &mut raw) });
        CudaBootstrap { raw, loaded_key: std::sync::Mutex::new(0) }
    }

    /// One engine over several GPUs: batches are sharded, the key is NCCL-broadcast at load time.
    pub fn with_devices(devices: &[i32]) -> Self {
        let p = TfheParams {
            n: params::tlwe_lv0::N as u32, big_n: params::trgsw_lv1::N as u32,
            l: params::trgsw_lv1::L as u32, bgbit: params::trgsw_lv1::BGBIT,
            basebit: params::trgsw_lv1::BASEBIT as u32, iks_t: params::trgsw_lv1::IKS_T as u32,
        };
        let mut raw = std::ptr::null_mut();
        check(unsafe { tfhe_engine_create_multi(&p, devices.as_ptr(), devices.len() as c_int, &mut raw) });
        CudaBootstrap { raw, loaded_key: std::sync::Mutex::new(0) }
    }

    /// Upload `ck` unless the device already holds a key with the same content.
    fn bind(&self, ck: &CloudKey) {
        let id = fingerprint(ck);
        let mut cur = self.loaded_key.lock().unwrap();
        if *cur != id {
            check(unsafe {
                tfhe_engine_load_cloud_key(self.raw, ck.decomposition_offset,
                    ck.blind_rotate_testvec.a.as_ptr(), ck.blind_rotate_testvec.b.as_ptr(),
                    ck.key_switching_key.as_ptr() as *const u32,      // Vec<TLWELv0{p:[u32;n+1]}>
                    ck.bootstrapping_key.as_ptr() as *const c_double) // Vec<TRGSWLv1FFT> = f64[n][2l][2][1024]
            });
            *cur = id;
        }
    }

    pub fn batch_gate(&self, op: Gate, inputs: &[(Ciphertext, Ciphertext)], ck: &CloudKey) -> Vec<Ciphertext> {
        self.bind(ck);
        let mut out = vec![Ciphertext::new(); inputs.len()];
        check(unsafe { tfhe_batch_gate(self.raw, op as c_int, inputs.as_ptr() as *const u32,
                                       out.as_mut_ptr() as *mut u32, inputs.len()) });
        out
    }

    pub fn batch_blind_rotate(&self, srcs: &[Ciphertext], ck: &CloudKey) -> Vec<TRLWELv1> {
        self.bind(ck);
        let mut out = vec![TRLWELv1::new(); srcs.len()];
        check(unsafe { tfhe_batch_blind_rotate(self.raw, srcs.as_ptr() as *const u32,
                                               out.as_mut_ptr() as *mut u32, srcs.len()) });
        out
    }

    /// LutBootstrap::bootstrap_func over a batch: the closure is tabulated on the host.
    pub fn batch_bootstrap_func<F: Fn(usize) -> usize>(&self, cts: &[Ciphertext], f: F,
                                                       message_modulus: usize, ck: &CloudKey) -> Vec<Ciphertext> {
        self.bind(ck);
        let table: Vec<u32> = (0..message_modulus).map(|x| (f(x) % message_modulus) as u32).collect();
        let mut out = vec![Ciphertext::new(); cts.len()];
        // the table goes into the engine's scratch slot: nothing to release, callable without bound
        check(unsafe { tfhe_batch_bootstrap_func(self.raw, table.as_ptr(), message_modulus as u32, 0.0,
                                                 cts.as_ptr() as *const u32, out.as_mut_ptr() as *mut u32, cts.len()) });
        out
    }

    /// Generator::generate_lookup_table kept on the device; released when the handle drops.
    pub fn lookup_table<F: Fn(usize) -> usize>(&self, f: F, message_modulus: usize, ck: &CloudKey) -> DeviceLut<'_> {
        self.bind(ck);
        let table: Vec<u32> = (0..message_modulus).map(|x| (f(x) % message_modulus) as u32).collect();
        let mut id: c_int = -1;
        check(unsafe { tfhe_lut_generate(self.raw, table.as_ptr(), message_modulus as u32, 0.0,
                                         std::ptr::null_mut(), &mut id) });
        DeviceLut { engine: self, id }
    }

    /// LutBootstrap::bootstrap_lut over a batch.
    pub fn batch_bootstrap_lut(&self, cts: &[Ciphertext], lut: &DeviceLut<'_>) -> Vec<Ciphertext> {
        let mut out = vec![Ciphertext::new(); cts.len()];
        check(unsafe { tfhe_batch_bootstrap_lut(self.raw, lut.id, cts.as_ptr() as *const u32,
                                                out.as_mut_ptr() as *mut u32, cts.len()) });
        out
    }

    /// FFTProcessor::batch_ifft / batch_fft / poly_mul (src/fft/mod.rs:80-107) on the device.
    pub fn batch_ifft(&self, inputs: &[[params::Torus; 1024]]) -> Vec<[f64; 1024]> {
        let mut out = vec![[0.0f64; 1024]; inputs.len()];
        check(unsafe { tfhe_batch_ifft(self.raw, inputs.as_ptr() as *const u32, out.as_mut_ptr() as *mut c_double, inputs.len()) });
        out
    }
    pub fn batch_fft(&self, inputs: &[[f64; 1024]]) -> Vec<[params::Torus; 1024]> {
        let mut out = vec![[0u32; 1024]; inputs.len()];
        check(unsafe { tfhe_batch_fft(self.raw, inputs.as_ptr() as *const c_double, out.as_mut_ptr() as *mut u32, inputs.len()) });
        out
    }
    pub fn batch_poly_mul(&self, a: &[[params::Torus; 1024]], b: &[[params::Torus; 1024]]) -> Vec<[params::Torus; 1024]> {
        assert_eq!(a.len(), b.len());
        let mut out = vec![[0u32; 1024]; a.len()];
        check(unsafe { tfhe_batch_poly_mul(self.raw, a.as_ptr() as *const u32, b.as_ptr() as *const u32,
                                           out.as_mut_ptr() as *mut u32, a.len()) });
        out
    }
}

/// A lookup table resident on the device (lut::LookupTable); its slot is returned on drop.
pub struct DeviceLut<'a> { engine: &'a CudaBootstrap, id: c_int }
impl Drop for DeviceLut<'_> {
    fn drop(&mut self) { unsafe { tfhe_lut_release(self.engine.raw, self.id); } }
}

impl CudaBootstrap {
    /// CloudKey::new(&secret_key) (src/key.rs:59-66) generated on the GPU; the key never
    /// exists in the reference layout on the host.  `seed == 0`: generator keyed from OS entropy (the
    /// only setting for real keys); any other value gives a reproducible, insecure test key.
    pub fn generate_cloud_key(&self, sk: &crate::key::SecretKey, seed: u64) {
        check(unsafe { tfhe_engine_generate_cloud_key(self.raw, sk.key_lv0.as_ptr(), sk.key_lv1.as_ptr(),
                                                      params::KSK_ALPHA, params::BSK_ALPHA, seed) });
        *self.loaded_key.lock().unwrap() = 2; // device-resident key, not tied to a CloudKey (fingerprints are odd)
    }

    /// proxy_reenc::reencrypt_tlwe_lv0 (src/proxy_reenc.rs:468-511) over a batch.
    pub fn batch_reencrypt(&self, key: &crate::proxy_reenc::ProxyReencryptionKey, cts: &[Ciphertext]) -> Vec<Ciphertext> {
        let mut h: *mut c_void = std::ptr::null_mut();
        check(unsafe { tfhe_reenc_key_load(self.raw, key.key_encryptions.as_ptr() as *const u32,
                                           key.base as u32, key.t as u32, &mut h) });
        let mut out = vec![Ciphertext::new(); cts.len()];
        check(unsafe { tfhe_batch_reencrypt(self.raw, h, cts.as_ptr() as *const u32,
                                            out.as_mut_ptr() as *mut u32, cts.len()) });
        unsafe { tfhe_reenc_key_destroy(h) };
        out
    }
}

impl Drop for CudaBootstrap {
    fn drop(&mut self) { unsafe { tfhe_engine_destroy(self.raw) } }
}

impl Bootstrap for CudaBootstrap {
    fn bootstrap(&self, ctxt: &Ciphertext, cloud_key: &CloudKey) -> Ciphertext {
        self.bind(cloud_key);
        let mut out = Ciphertext::new();
        check(unsafe { tfhe_batch_bootstrap(self.raw, ctxt.p.as_ptr(), out.p.as_mut_ptr(), 1, 1) });
        out
    }
    fn bootstrap_without_key_switch(&self, ctxt: &Ciphertext, cloud_key: &CloudKey) -> Ciphertext {
        self.bind(cloud_key);
        let mut out = Ciphertext::new();
        check(unsafe { tfhe_batch_bootstrap(self.raw, ctxt.p.as_ptr(), out.p.as_mut_ptr(), 1, 0) });
        out
    }
    fn name(&self) -> &str { "cuda-b200" }
}

// gates.rs:352-547 re-routed: `pub fn batch_nand(inputs, cloud_key)` becomes
//   cuda::default_engine().batch_gate(Gate::Nand, inputs, cloud_key)
pub fn default_engine() -> &'static CudaBootstrap {
    static ENGINE: std::sync::OnceLock<CudaBootstrap> = std::sync::OnceLock::new();
    ENGINE.get_or_init(|| CudaBootstrap::new(0))
}
