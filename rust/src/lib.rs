//! jpeg-encoder's public API (`Encoder::new` / setters / `encode` / `encode_image`) forwarding to the B200 C ABI
//! (include/jpegenc_b200.h). Source only: this repository's image has no Rust toolchain, so the crate is checked
//! textually against the reference's public items (tests/test_rust_shim.py), not compiled.
//!
//! Public surface = /root/reference/src/lib.rs:45-49:
//!   encoder::{ColorType, Encoder, JpegColorType, SamplingFactor}, error::EncodingError,
//!   image_buffer::{ImageBuffer, cmyk_to_ycck, rgb_to_ycbcr}, quantization::QuantizationTableType,
//!   writer::{JfifWrite, PixelDensity, PixelDensityUnit}.
//! `EncodingError` has exactly the reference's variants (src/error.rs:6-28): a CUDA failure is reported as
//! `IoError(std::io::Error)` of kind `Other`, so user code that matches exhaustively keeps compiling.
use std::fmt::Display;
use std::error::Error;
use std::fs::File;
use std::io::BufWriter;
use std::os::raw::{c_char, c_int, c_void};
use std::path::Path;

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
    fn jpgb_last_error(enc: *const JpgbEncoder) -> *const c_char;
    fn jpgb_encode_to_sink(enc: *mut JpgbEncoder, p: *const JpgbParams, pixels: *const u8, len: usize,
                           write_all: WriteAll, user: *mut c_void) -> c_int;
    fn jpgb_encode_planar_to_sink(enc: *mut JpgbEncoder, p: *const JpgbParams, planes: *const *const u8, plane_len: usize,
                                  write_all: WriteAll, user: *mut c_void) -> c_int;
}

// ---- src/encoder.rs:21-65 ------------------------------------------------------------------------
/// # Color types used in encoding
#[derive(Copy, Clone, Debug, Eq, PartialEq)]
pub enum JpegColorType {
    /// One component grayscale colorspace
    Luma,
    /// Three component YCbCr colorspace
    Ycbcr,
    /// 4 Component CMYK colorspace
    Cmyk,
    /// 4 Component YCbCrK colorspace
    Ycck,
}

impl JpegColorType {
    pub(crate) fn get_num_components(self) -> usize {
        match self { JpegColorType::Luma => 1, JpegColorType::Ycbcr => 3, JpegColorType::Cmyk | JpegColorType::Ycck => 4 }
    }
    /// the C ABI's colour code of the matching verbatim input type (JPGB_LUMA / YCBCR / CMYK / YCCK)
    fn abi_code(self) -> u8 { match self { JpegColorType::Luma => 0, JpegColorType::Ycbcr => 5, JpegColorType::Cmyk => 6, JpegColorType::Ycck => 8 } }
}

// ---- src/encoder.rs:72-111 -----------------------------------------------------------------------
/// # Color types for input images
#[derive(Copy, Clone, Debug, Eq, PartialEq)]
pub enum ColorType {
    /// Grayscale with 1 byte per pixel
    Luma,
    /// RGB with 3 bytes per pixel
    Rgb,
    /// Red, Green, Blue with 4 bytes per pixel. The alpha channel will be ignored during encoding.
    Rgba,
    /// RGB with 3 bytes per pixel
    Bgr,
    /// RGBA with 4 bytes per pixel. The alpha channel will be ignored during encoding.
    Bgra,
    /// YCbCr with 3 bytes per pixel.
    Ycbcr,
    /// CMYK with 4 bytes per pixel.
    Cmyk,
    /// CMYK with 4 bytes per pixel. Encoded as YCCK (YCbCrK)
    CmykAsYcck,
    /// YCCK (YCbCrK) with 4 bytes per pixel.
    Ycck,
}

impl ColorType {
    pub(crate) fn get_bytes_per_pixel(self) -> usize {
        use ColorType::*;
        match self { Luma => 1, Rgb | Bgr | Ycbcr => 3, Rgba | Bgra | Cmyk | CmykAsYcck | Ycck => 4 }
    }
}

// ---- src/encoder.rs:113-187 ----------------------------------------------------------------------
#[repr(u8)]
#[derive(Copy, Clone, Debug, Eq, PartialEq)]
#[allow(non_camel_case_types)]
pub enum SamplingFactor {
    F_1_1 = 1 << 4 | 1, F_2_1 = 2 << 4 | 1, F_1_2 = 1 << 4 | 2, F_2_2 = 2 << 4 | 2,
    F_4_1 = 4 << 4 | 1, F_4_2 = 4 << 4 | 2, F_1_4 = 1 << 4 | 4, F_2_4 = 2 << 4 | 4,
    /// Alias for F_1_1
    R_4_4_4 = 0x80 | 1 << 4 | 1,
    /// Alias for F_1_2
    R_4_4_0 = 0x80 | 1 << 4 | 2,
    /// Alias for F_1_4
    R_4_4_1 = 0x80 | 1 << 4 | 4,
    /// Alias for F_2_1
    R_4_2_2 = 0x80 | 2 << 4 | 1,
    /// Alias for F_2_2
    R_4_2_0 = 0x80 | 2 << 4 | 2,
    /// Alias for F_2_4
    R_4_2_1 = 0x80 | 2 << 4 | 4,
    /// Alias for F_4_1
    R_4_1_1 = 0x80 | 4 << 4 | 1,
    /// Alias for F_4_2
    R_4_1_0 = 0x80 | 4 << 4 | 2,
}

impl SamplingFactor {
    /// Get variant for supplied factors or None if not supported
    pub fn from_factors(horizontal: u8, vertical: u8) -> Option<SamplingFactor> {
        use SamplingFactor::*;
        match (horizontal, vertical) {
            (1, 1) => Some(F_1_1), (1, 2) => Some(F_1_2), (1, 4) => Some(F_1_4),
            (2, 1) => Some(F_2_1), (2, 2) => Some(F_2_2), (2, 4) => Some(F_2_4),
            (4, 1) => Some(F_4_1), (4, 2) => Some(F_4_2),
            _ => None,
        }
    }
    pub(crate) fn get_sampling_factors(self) -> (u8, u8) {
        let value = self as u8;
        ((value >> 4) & 0x07, value & 0xf)
    }
}

// ---- src/quantization.rs:8-40 --------------------------------------------------------------------
/// # Quantization table used for encoding
#[derive(Debug, Clone)]
pub enum QuantizationTableType {
    /// Sample quantization tables given in Annex K (Clause K.1) of Recommendation ITU-T T.81 (1992) | ISO/IEC 10918-1:1994.
    Default,
    /// Flat
    Flat,
    /// Custom, tuned for MS-SSIM
    CustomMsSsim,
    /// Custom, tuned for PSNR-HVS
    CustomPsnrHvs,
    /// ImageMagick table by N. Robidoux
    ImageMagick,
    /// Relevance of human vision to JPEG-DCT compression (1992) Klein, Silverstein and Carney.
    KleinSilversteinCarney,
    /// DCTune perceptual optimization of compressed dental X-Rays (1997) Watson, Taylor, Borthwick
    DentalXRays,
    /// A visual detection model for DCT coefficient quantization (12/9/93) Ahumada, Watson, Peterson
    VisualDetectionModel,
    /// An improved detection model for DCT coefficient quantization (1993) Peterson, Ahumada and Watson
    ImprovedDetectionModel,
    /// A user supplied quantization table
    Custom(Box<[u16; 64]>),
}

impl QuantizationTableType {
    fn kind(&self) -> u8 {
        use QuantizationTableType::*;
        match self { Default => 0, Flat => 1, CustomMsSsim => 2, CustomPsnrHvs => 3, ImageMagick => 4,
                     KleinSilversteinCarney => 5, DentalXRays => 6, VisualDetectionModel => 7, ImprovedDetectionModel => 8, Custom(_) => 9 }
    }
}

// ---- src/writer.rs:16-106 ------------------------------------------------------------------------
/// Represents the pixel density of an image
#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub struct PixelDensity {
    /// A couple of values for (Xdensity, Ydensity)
    pub density: (u16, u16),
    /// The unit in which the density is measured
    pub unit: PixelDensityUnit,
}

impl PixelDensity {
    /// Creates the most common pixel density type:
    /// the horizontal and the vertical density are equal,
    /// and measured in pixels per inch.
    #[must_use]
    pub fn dpi(density: u16) -> Self {
        PixelDensity { density: (density, density), unit: PixelDensityUnit::Inches }
    }
}

impl Default for PixelDensity {
    /// Returns a pixel density with a pixel aspect ratio of 1
    fn default() -> Self {
        PixelDensity { density: (1, 1), unit: PixelDensityUnit::PixelAspectRatio }
    }
}

/// Represents a unit in which the density of an image is measured
#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub enum PixelDensityUnit {
    /// Represents the absence of a unit, the values indicate only a
    /// [pixel aspect ratio](https://en.wikipedia.org/wiki/Pixel_aspect_ratio)
    PixelAspectRatio,
    /// Pixels per inch (2.54 cm)
    Inches,
    /// Pixels per centimeter
    Centimeters,
}

/// # Buffered writer for JFIF data
pub trait JfifWrite {
    /// Writes the whole buffer. The behavior must be identical to std::io::Write::write_all
    fn write_all(&mut self, buf: &[u8]) -> Result<(), EncodingError>;
}

impl<W: std::io::Write + ?Sized> JfifWrite for W {
    #[inline(always)]
    fn write_all(&mut self, buf: &[u8]) -> Result<(), EncodingError> {
        self.write_all(buf)?;
        Ok(())
    }
}

// ---- src/error.rs:6-75 ---------------------------------------------------------------------------
/// # The error type for encoding
#[derive(Debug)]
pub enum EncodingError {
    /// An invalid app segment number has been used
    InvalidAppSegment(u8),
    /// App segment exceeds maximum allowed data length
    AppSegmentTooLarge(usize),
    /// Color profile exceeds maximum allowed data length
    IccTooLarge(usize),
    /// Image data is too short
    BadImageData { length: usize, required: usize },
    /// Width or height is zero
    ZeroImageDimensions { width: u16, height: u16 },
    /// An io error occurred during writing. Also carries device failures (kind `Other`, message "CUDA: ..."):
    /// there is no CPU fallback.
    IoError(std::io::Error),
    /// An io error occurred during writing (Should be used in no_std cases instead of IoError)
    Write(String),
}

impl From<std::io::Error> for EncodingError {
    fn from(err: std::io::Error) -> EncodingError {
        EncodingError::IoError(err)
    }
}

impl Display for EncodingError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        use EncodingError::*;
        match self {
            InvalidAppSegment(nr) => write!(f, "Invalid app segment number: {}", nr),
            AppSegmentTooLarge(length) => write!(f, "App segment exceeds maximum allowed data length of 65533: {}", length),
            IccTooLarge(length) => write!(f, "ICC profile exceeds maximum allowed data length: {}", length),
            BadImageData { length, required } => write!(f, "Image data too small for dimensions and color_type: {} need at least {}", length, required),
            ZeroImageDimensions { width, height } => write!(f, "Image dimensions must be non zero: {}x{}", width, height),
            IoError(err) => err.fmt(f),
            Write(err) => write!(f, "{}", err),
        }
    }
}

impl Error for EncodingError {
    fn source(&self) -> Option<&(dyn Error + 'static)> {
        match self {
            EncodingError::IoError(err) => Some(err),
            _ => None,
        }
    }
}

// ---- src/image_buffer.rs:9-38, 86-98 -------------------------------------------------------------
/// Conversion from RGB to YCbCr. Host-side helper for `ImageBuffer` implementors (the reference exports it
/// for exactly that, src/image_buffer.rs:57-84); `Encoder::encode` converts on the GPU with the same integers.
#[inline]
pub fn rgb_to_ycbcr(r: u8, g: u8, b: u8) -> (u8, u8, u8) {
    let (r, g, b) = (r as i32, g as i32, b as i32);
    let y = 19595 * r + 38470 * g + 7471 * b;
    let cb = -11059 * r - 21709 * g + 32768 * b + (128 << 16);
    let cr = 32768 * r - 27439 * g - 5329 * b + (128 << 16);
    (((y + 0x7FFF) >> 16) as u8, ((cb + 0x7FFF) >> 16) as u8, ((cr + 0x7FFF) >> 16) as u8)
}

/// Conversion from CMYK to YCCK (YCbCrK)
#[inline]
pub fn cmyk_to_ycck(c: u8, m: u8, y: u8, k: u8) -> (u8, u8, u8, u8) {
    let (y, cb, cr) = rgb_to_ycbcr(c, m, y);
    (y, cb, cr, 255 - k)
}

/// # Buffer used as input value for image encoding
pub trait ImageBuffer {
    /// The color type used in the image encoding
    fn get_jpeg_color_type(&self) -> JpegColorType;
    /// Width of the image
    fn width(&self) -> u16;
    /// Height of the image
    fn height(&self) -> u16;
    /// Add color values for the row to color component buffers
    fn fill_buffers(&self, y: u16, buffers: &mut [Vec<u8>; 4]);
}

// ---- src/encoder.rs:213-515 ----------------------------------------------------------------------
/// # The JPEG encoder
pub struct Encoder<W: JfifWrite> {
    w: W, quality: u8, density: PixelDensity, tables: [QuantizationTableType; 2], sampling: SamplingFactor,
    progressive_scans: Option<u8>, restart_interval: Option<u16>, optimize: bool, apps: Vec<(u8, Vec<u8>)>,
}

struct SinkState<'a, W: JfifWrite> { w: &'a mut W, err: Option<EncodingError> }
unsafe extern "C" fn sink_trampoline<W: JfifWrite>(user: *mut c_void, buf: *const u8, len: usize) -> c_int {
    let st = &mut *(user as *mut SinkState<W>);
    match st.w.write_all(std::slice::from_raw_parts(buf, len)) { Ok(()) => 0, Err(e) => { st.err = Some(e); 1 } }
}

/// One encoder context (CUDA stream + device buffers) per host thread, destroyed with the thread.
struct Context(*mut JpgbEncoder);
impl Drop for Context { fn drop(&mut self) { if !self.0.is_null() { unsafe { jpgb_encoder_destroy(self.0) } } } }

fn device_error(ctx: *const JpgbEncoder, rc: c_int) -> EncodingError {
    let msg = unsafe {
        let p = jpgb_last_error(ctx);
        if p.is_null() { String::new() } else { std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned() }
    };
    EncodingError::IoError(std::io::Error::new(std::io::ErrorKind::Other, format!("CUDA: jpegenc_b200 error {}: {}", rc, msg)))
}

impl<W: JfifWrite> Encoder<W> {
    /// Create a new encoder with the given quality
    pub fn new(w: W, quality: u8) -> Encoder<W> {
        Encoder { w, quality, density: PixelDensity::default(),
                  tables: [QuantizationTableType::Default, QuantizationTableType::Default],
                  sampling: if quality < 90 { SamplingFactor::F_2_2 } else { SamplingFactor::F_1_1 },
                  progressive_scans: None, restart_interval: None, optimize: false, apps: Vec::new() }
    }
    pub fn set_density(&mut self, density: PixelDensity) { self.density = density; }
    pub fn density(&self) -> PixelDensity { self.density }
    pub fn set_sampling_factor(&mut self, sampling: SamplingFactor) { self.sampling = sampling; }
    pub fn sampling_factor(&self) -> SamplingFactor { self.sampling }
    pub fn set_quantization_tables(&mut self, luma: QuantizationTableType, chroma: QuantizationTableType) { self.tables = [luma, chroma]; }
    pub fn quantization_tables(&self) -> &[QuantizationTableType; 2] { &self.tables }
    pub fn set_progressive(&mut self, progressive: bool) { self.progressive_scans = if progressive { Some(4) } else { None }; }
    pub fn set_progressive_scans(&mut self, scans: u8) {
        assert!((2..=64).contains(&scans), "Invalid number of scans: {}", scans);
        self.progressive_scans = Some(scans);
    }
    pub fn progressive_scans(&self) -> Option<u8> { self.progressive_scans }
    pub fn set_restart_interval(&mut self, interval: u16) { self.restart_interval = if interval == 0 { None } else { Some(interval) }; }
    pub fn restart_interval(&self) -> Option<u16> { self.restart_interval }
    pub fn set_optimized_huffman_tables(&mut self, optimize_huffman_table: bool) { self.optimize = optimize_huffman_table; }
    pub fn optimized_huffman_tables(&self) -> bool { self.optimize }
    pub fn add_app_segment(&mut self, segment_nr: u8, data: Vec<u8>) -> Result<(), EncodingError> {
        if segment_nr == 0 || segment_nr > 15 { Err(EncodingError::InvalidAppSegment(segment_nr)) }
        else if data.len() > 65533 { Err(EncodingError::AppSegmentTooLarge(data.len())) }
        else { self.apps.push((segment_nr, data)); Ok(()) }
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

    fn context() -> Result<*mut JpgbEncoder, EncodingError> {
        thread_local! { static CTX: std::cell::RefCell<Context> = std::cell::RefCell::new(Context(std::ptr::null_mut())); }
        CTX.with(|c| {
            let mut c = c.borrow_mut();
            if c.0.is_null() {
                let mut e = std::ptr::null_mut();
                let rc = unsafe { jpgb_encoder_create(0, std::ptr::null_mut(), &mut e) };
                if rc != 0 { return Err(device_error(std::ptr::null(), rc)); }
                c.0 = e;
            }
            Ok(c.0)
        })
    }

    /// Encode an image: same contract as the reference (length check first, then zero dimensions); the work
    /// runs on the B200 and the file is handed to `W::write_all` straight from the context's pinned buffer.
    pub fn encode(mut self, data: &[u8], width: u16, height: u16, color_type: ColorType) -> Result<(), EncodingError> {
        let required = width as usize * height as usize * color_type.get_bytes_per_pixel();
        if data.len() < required { return Err(EncodingError::BadImageData { length: data.len(), required }); }
        if width == 0 || height == 0 { return Err(EncodingError::ZeroImageDimensions { width, height }); }
        let apps: Vec<JpgbApp> = self.apps.iter().map(|(nr, d)| JpgbApp { nr: *nr, data: d.as_ptr(), len: d.len() as u32 }).collect();
        unsafe {
            let p = self.params(width, height, color_type as u8, &apps);
            let ctx = Self::context()?;
            let mut st = SinkState { w: &mut self.w, err: None };
            let rc = jpgb_encode_to_sink(ctx, &p, data.as_ptr(), data.len(), sink_trampoline::<W>, &mut st as *mut _ as *mut c_void);
            if let Some(e) = st.err { return Err(e); }
            if rc != 0 { return Err(device_error(ctx, rc)); }
        }
        Ok(())
    }

    /// Encode an image: the user's `fill_buffers` runs on the host, row by row, exactly as the reference calls
    /// it; the GPU takes the resulting planes as already-converted component samples and does padding,
    /// decimation, fDCT, quantization and entropy coding.
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
            let mut st = SinkState { w: &mut self.w, err: None };
            let rc = jpgb_encode_planar_to_sink(ctx, &p, ptrs.as_ptr(), plane_len, sink_trampoline::<W>, &mut st as *mut _ as *mut c_void);
            if let Some(e) = st.err { return Err(e); }
            if rc != 0 { return Err(device_error(ctx, rc)); }
        }
        Ok(())
    }
}

impl Encoder<BufWriter<File>> {
    /// Create a new decoder that writes into a file
    pub fn new_file<P: AsRef<Path>>(path: P, quality: u8) -> Result<Encoder<BufWriter<File>>, EncodingError> {
        let file = File::create(path)?;
        let buf = BufWriter::new(file);
        Ok(Self::new(buf, quality))
    }
}
