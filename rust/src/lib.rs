//! jpeg-encoder public API (Encoder::new / setters / encode) forwarding to the B200 C ABI
//! (include/jpegenc_b200.h). Source only: this repository's image has no Rust toolchain.
//! Mirrors /root/reference/src/encoder.rs:213-515, src/writer.rs:16-106, src/error.rs.
use std::os::raw::{c_int, c_void};

#[repr(C)]
struct JpgbApp { nr: u8, data: *const u8, len: u32 }

#[repr(C)]
struct JpgbParams {
    width: u16, height: u16,
    color_type: u8, quality: u8, sampling: u8,
    qtable_kind: [u8; 2],
    qtable_custom: [[u16; 64]; 2],
    progressive_scans: u8, optimize_huffman: u8,
    restart_interval: u16,
    density_unit: u8, density_x: u16, density_y: u16,
    n_app: u32, apps: *const JpgbApp,
}

#[repr(C)] struct JpgbEncoder { _private: [u8; 0] }
type WriteAll = unsafe extern "C" fn(user: *mut c_void, buf: *const u8, len: usize) -> c_int;

extern "C" {
    fn jpgb_params_default(p: *mut JpgbParams, quality: u8);
    fn jpgb_encoder_create(device: c_int, stream: *mut c_void, out: *mut *mut JpgbEncoder) -> c_int;
    fn jpgb_encoder_destroy(enc: *mut JpgbEncoder);
    fn jpgb_encode_to_sink(enc: *mut JpgbEncoder, p: *const JpgbParams, pixels: *const u8, len: usize,
                           write_all: WriteAll, user: *mut c_void) -> c_int;
    fn jpgb_encode_planar(enc: *mut JpgbEncoder, p: *const JpgbParams, planes: *const *const u8, plane_len: usize,
                          out: *mut *mut u8, out_len: *mut usize) -> c_int;
    fn jpgb_free(buf: *mut c_void);
}

/// src/encoder.rs:27-65
#[derive(Copy, Clone, Debug, Eq, PartialEq)]
pub enum JpegColorType { Luma, Ycbcr, Cmyk, Ycck }
impl JpegColorType {
    fn get_num_components(self) -> usize { match self { JpegColorType::Luma => 1, JpegColorType::Ycbcr => 3, _ => 4 } }
    /// the C ABI's colour code of the matching verbatim input type (JPGB_LUMA / YCBCR / CMYK / YCCK)
    fn abi_code(self) -> u8 { match self { JpegColorType::Luma => 0, JpegColorType::Ycbcr => 5, JpegColorType::Cmyk => 6, JpegColorType::Ycck => 8 } }
}

/// src/image_buffer.rs:86-98: a user pixel format. `fill_buffers` appends one row of converted samples per component.
pub trait ImageBuffer {
    fn get_jpeg_color_type(&self) -> JpegColorType;
    fn width(&self) -> u16;
    fn height(&self) -> u16;
    fn fill_buffers(&self, y: u16, buffers: &mut [Vec<u8>; 4]);
}

#[derive(Copy, Clone, Debug, Eq, PartialEq)]
pub enum ColorType { Luma, Rgb, Rgba, Bgr, Bgra, Ycbcr, Cmyk, CmykAsYcck, Ycck }
impl ColorType {
    fn bpp(self) -> usize { match self { ColorType::Luma => 1, ColorType::Rgb | ColorType::Bgr | ColorType::Ycbcr => 3, _ => 4 } }
}

#[repr(u8)] #[derive(Copy, Clone, Debug, Eq, PartialEq)] #[allow(non_camel_case_types)]
pub enum SamplingFactor {
    F_1_1 = 1 << 4 | 1, F_2_1 = 2 << 4 | 1, F_1_2 = 1 << 4 | 2, F_2_2 = 2 << 4 | 2,
    F_4_1 = 4 << 4 | 1, F_4_2 = 4 << 4 | 2, F_1_4 = 1 << 4 | 4, F_2_4 = 2 << 4 | 4,
    R_4_4_4 = 0x80 | 1 << 4 | 1, R_4_4_0 = 0x80 | 1 << 4 | 2, R_4_4_1 = 0x80 | 1 << 4 | 4, R_4_2_2 = 0x80 | 2 << 4 | 1,
    R_4_2_0 = 0x80 | 2 << 4 | 2, R_4_2_1 = 0x80 | 2 << 4 | 4, R_4_1_1 = 0x80 | 4 << 4 | 1, R_4_1_0 = 0x80 | 4 << 4 | 2,
}

#[derive(Debug, Clone)]
pub enum QuantizationTableType {
    Default, Flat, CustomMsSsim, CustomPsnrHvs, ImageMagick, KleinSilversteinCarney, DentalXRays,
    VisualDetectionModel, ImprovedDetectionModel, Custom(Box<[u16; 64]>),
}
impl QuantizationTableType {
    fn kind(&self) -> u8 {
        use QuantizationTableType::*;
        match self { Default => 0, Flat => 1, CustomMsSsim => 2, CustomPsnrHvs => 3, ImageMagick => 4,
                     KleinSilversteinCarney => 5, DentalXRays => 6, VisualDetectionModel => 7, ImprovedDetectionModel => 8, Custom(_) => 9 }
    }
}

#[derive(Clone, Copy, Debug, Eq, PartialEq)] pub enum PixelDensityUnit { PixelAspectRatio, Inches, Centimeters }
#[derive(Clone, Copy, Debug, Eq, PartialEq)] pub struct PixelDensity { pub density: (u16, u16), pub unit: PixelDensityUnit }
impl PixelDensity { pub fn dpi(d: u16) -> Self { PixelDensity { density: (d, d), unit: PixelDensityUnit::Inches } } }
impl Default for PixelDensity { fn default() -> Self { PixelDensity { density: (1, 1), unit: PixelDensityUnit::PixelAspectRatio } } }

#[derive(Debug)]
pub enum EncodingError {
    InvalidAppSegment(u8), AppSegmentTooLarge(usize), IccTooLarge(usize),
    BadImageData { length: usize, required: usize }, ZeroImageDimensions { width: u16, height: u16 },
    IoError(std::io::Error), Write(String),
    /// no usable B200, or a CUDA call failed (there is no CPU fallback)
    Cuda(i32),
}

pub trait JfifWrite { fn write_all(&mut self, buf: &[u8]) -> Result<(), EncodingError>; }
impl<W: std::io::Write + ?Sized> JfifWrite for W {
    fn write_all(&mut self, buf: &[u8]) -> Result<(), EncodingError> { std::io::Write::write_all(self, buf).map_err(EncodingError::IoError) }
}

pub struct Encoder<W: JfifWrite> {
    w: W, quality: u8, density: PixelDensity, tables: [QuantizationTableType; 2], sampling: SamplingFactor,
    progressive_scans: Option<u8>, restart_interval: Option<u16>, optimize: bool, apps: Vec<(u8, Vec<u8>)>,
}

struct SinkState<'a, W: JfifWrite> { w: &'a mut W, err: Option<EncodingError> }
unsafe extern "C" fn sink_trampoline<W: JfifWrite>(user: *mut c_void, buf: *const u8, len: usize) -> c_int {
    let st = &mut *(user as *mut SinkState<W>);
    match st.w.write_all(std::slice::from_raw_parts(buf, len)) { Ok(()) => 0, Err(e) => { st.err = Some(e); 1 } }
}

impl<W: JfifWrite> Encoder<W> {
    pub fn new(w: W, quality: u8) -> Encoder<W> {
        Encoder { w, quality, density: PixelDensity::default(),
                  tables: [QuantizationTableType::Default, QuantizationTableType::Default],
                  sampling: if quality < 90 { SamplingFactor::F_2_2 } else { SamplingFactor::F_1_1 },
                  progressive_scans: None, restart_interval: None, optimize: false, apps: Vec::new() }
    }
    pub fn set_density(&mut self, d: PixelDensity) { self.density = d; }
    pub fn density(&self) -> PixelDensity { self.density }
    pub fn set_sampling_factor(&mut self, s: SamplingFactor) { self.sampling = s; }
    pub fn sampling_factor(&self) -> SamplingFactor { self.sampling }
    pub fn set_quantization_tables(&mut self, luma: QuantizationTableType, chroma: QuantizationTableType) { self.tables = [luma, chroma]; }
    pub fn quantization_tables(&self) -> &[QuantizationTableType; 2] { &self.tables }
    pub fn set_progressive(&mut self, p: bool) { self.progressive_scans = if p { Some(4) } else { None }; }
    pub fn set_progressive_scans(&mut self, scans: u8) { assert!((2..=64).contains(&scans), "Invalid number of scans: {}", scans); self.progressive_scans = Some(scans); }
    pub fn progressive_scans(&self) -> Option<u8> { self.progressive_scans }
    pub fn set_restart_interval(&mut self, i: u16) { self.restart_interval = if i == 0 { None } else { Some(i) }; }
    pub fn restart_interval(&self) -> Option<u16> { self.restart_interval }
    pub fn set_optimized_huffman_tables(&mut self, o: bool) { self.optimize = o; }
    pub fn optimized_huffman_tables(&self) -> bool { self.optimize }
    pub fn add_app_segment(&mut self, nr: u8, data: Vec<u8>) -> Result<(), EncodingError> {
        if nr == 0 || nr > 15 { Err(EncodingError::InvalidAppSegment(nr)) }
        else if data.len() > 65533 { Err(EncodingError::AppSegmentTooLarge(data.len())) }
        else { self.apps.push((nr, data)); Ok(()) }
    }
    pub fn add_icc_profile(&mut self, data: &[u8]) -> Result<(), EncodingError> {
        const MARKER: &[u8; 12] = b"ICC_PROFILE\0";
        const MAX: usize = 65535 - 2 - 12 - 2;
        let n = (data.len() + MAX - 1) / MAX;
        if n >= 255 { return Err(EncodingError::IccTooLarge(data.len())); }
        for (i, c) in data.chunks(MAX).enumerate() {
            let mut v = Vec::with_capacity(MAX); v.extend_from_slice(MARKER); v.push(i as u8 + 1); v.push(n as u8); v.extend_from_slice(c);
            self.add_app_segment(2, v)?;
        }
        Ok(())
    }
    pub fn add_exif_metadata(&mut self, data: &[u8]) -> Result<(), EncodingError> {
        let mut f = vec![0x45, 0x78, 0x69, 0x66, 0x00, 0x00]; f.extend_from_slice(data); self.add_app_segment(1, f)
    }

    /// The state of this encoder as the C ABI takes it (`apps` must outlive the returned struct).
    unsafe fn params(&self, width: u16, height: u16, color_code: u8, apps: &[JpgbApp]) -> JpgbParams {
        let mut p: JpgbParams = std::mem::zeroed();
        jpgb_params_default(&mut p, self.quality);
        p.width = width; p.height = height; p.color_type = color_code; p.sampling = self.sampling as u8;
        for i in 0..2 {
            p.qtable_kind[i] = self.tables[i].kind();
            if let QuantizationTableType::Custom(t) = &self.tables[i] { p.qtable_custom[i] = **t; }
        }
        p.progressive_scans = self.progressive_scans.unwrap_or(0);
        p.optimize_huffman = self.optimize as u8;
        p.restart_interval = self.restart_interval.unwrap_or(0);
        p.density_unit = match self.density.unit { PixelDensityUnit::PixelAspectRatio => 0, PixelDensityUnit::Inches => 1, PixelDensityUnit::Centimeters => 2 };
        p.density_x = self.density.density.0; p.density_y = self.density.density.1;
        p.n_app = apps.len() as u32; p.apps = apps.as_ptr();
        p
    }

    /// One encoder context (CUDA stream + device buffers) per host thread; there is no CPU fallback.
    unsafe fn context() -> Result<*mut JpgbEncoder, EncodingError> {
        thread_local! { static CTX: std::cell::Cell<*mut JpgbEncoder> = std::cell::Cell::new(std::ptr::null_mut()); }
        let ctx = CTX.with(|c| { if c.get().is_null() { let mut e = std::ptr::null_mut(); if jpgb_encoder_create(0, std::ptr::null_mut(), &mut e) == 0 { c.set(e); } } c.get() });
        if ctx.is_null() { Err(EncodingError::Cuda(8)) } else { Ok(ctx) }
    }

    /// Encoder::encode (src/encoder.rs:440-503): same contract as the reference; the work runs on the B200.
    pub fn encode(mut self, data: &[u8], width: u16, height: u16, color_type: ColorType) -> Result<(), EncodingError> {
        let required = width as usize * height as usize * color_type.bpp();
        if data.len() < required { return Err(EncodingError::BadImageData { length: data.len(), required }); }
        if width == 0 || height == 0 { return Err(EncodingError::ZeroImageDimensions { width, height }); }
        let apps: Vec<JpgbApp> = self.apps.iter().map(|(nr, d)| JpgbApp { nr: *nr, data: d.as_ptr(), len: d.len() as u32 }).collect();
        unsafe {
            let p = self.params(width, height, color_type as u8, &apps);
            let ctx = Self::context()?;
            let mut st = SinkState { w: &mut self.w, err: None };
            let rc = jpgb_encode_to_sink(ctx, &p, data.as_ptr(), data.len(), sink_trampoline::<W>, &mut st as *mut _ as *mut c_void);
            if let Some(e) = st.err { return Err(e); }
            if rc != 0 { return Err(EncodingError::Cuda(rc)); }
        }
        Ok(())
    }

    /// Encoder::encode_image (src/encoder.rs:506-515): the user's `fill_buffers` runs on the host, row by row, exactly as
    /// the reference calls it; the GPU takes the resulting planes as already-converted component samples
    /// (jpgb_encode_planar) and does padding, decimation, fDCT, quantization and entropy coding.
    pub fn encode_image<I: ImageBuffer>(mut self, image: I) -> Result<(), EncodingError> {
        let (width, height) = (image.width(), image.height());
        if width == 0 || height == 0 { return Err(EncodingError::ZeroImageDimensions { width, height }); }
        let jct = image.get_jpeg_color_type();
        let mut planes: [Vec<u8>; 4] = Default::default();
        for y in 0..height { image.fill_buffers(y, &mut planes); }
        let plane_len = width as usize * height as usize;
        let n = jct.get_num_components();
        if planes[..n].iter().any(|pl| pl.len() < plane_len) {
            return Err(EncodingError::BadImageData { length: planes[..n].iter().map(|pl| pl.len()).min().unwrap_or(0), required: plane_len });
        }
        let ptrs: [*const u8; 4] = [planes[0].as_ptr(), planes[1].as_ptr(), planes[2].as_ptr(), planes[3].as_ptr()];
        let apps: Vec<JpgbApp> = self.apps.iter().map(|(nr, d)| JpgbApp { nr: *nr, data: d.as_ptr(), len: d.len() as u32 }).collect();
        unsafe {
            let p = self.params(width, height, jct.abi_code(), &apps);
            let ctx = Self::context()?;
            let (mut out, mut out_len) = (std::ptr::null_mut::<u8>(), 0usize);
            let rc = jpgb_encode_planar(ctx, &p, ptrs.as_ptr(), plane_len, &mut out, &mut out_len);
            if rc != 0 { return Err(EncodingError::Cuda(rc)); }
            let res = self.w.write_all(std::slice::from_raw_parts(out, out_len));
            jpgb_free(out as *mut c_void);
            res
        }
    }
}
#[allow(dead_code)] fn _keep(e: *mut JpgbEncoder) { unsafe { jpgb_encoder_destroy(e) } }
