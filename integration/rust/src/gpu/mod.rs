//! GPU-backed twins of the crate's automata.  Construction stays in the crate; the serialized automaton
//! crosses the FFI once; batches of haystacks are scanned by libdaachorse_b200 on a B200.
//! Not compiled in this repository (no Rust toolchain in the build image).
//!
//! Needs two one-line accessors in the crate, `pub(crate) fn match_kind(&self) -> MatchKind { self.match_kind }`
//! in src/bytewise.rs and src/charwise.rs (the field is private to those modules); `Match`'s fields and
//! `MatchKind::is_standard / is_leftmost` are visible here because this module hangs off the crate root.
pub mod ffi;

use crate::{CharwiseDoubleArrayAhoCorasick, DoubleArrayAhoCorasick, Match, MatchKind};
use core::ffi::CStr;

#[derive(Debug)]
pub struct GpuError {
    pub code: i32,
    pub message: String,
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::dach_last_error()) }.to_string_lossy().into_owned()
}

fn check(code: i32) -> Result<(), GpuError> {
    if code == ffi::DACH_OK {
        Ok(())
    } else {
        Err(GpuError { code, message: last_error() })
    }
}

/// Matches of a batch: `matches[offsets[i]..offsets[i + 1]]` belong to haystack i, in the order the
/// crate's iterator yields them.
pub struct BatchMatches {
    pub matches: Vec<Match<u32>>,
    pub offsets: Vec<u64>,
}

struct Device {
    dev: *mut ffi::DachDev,
}
impl Drop for Device {
    fn drop(&mut self) {
        unsafe { ffi::dach_dev_free(self.dev) }
    }
}

fn upload(bytes: &[u8], charwise: bool, device: i32) -> Result<Device, GpuError> {
    let (mut h, mut used, mut dev) = (core::ptr::null_mut(), 0usize, core::ptr::null_mut());
    check(unsafe { ffi::dach_pma_deserialize(bytes.as_ptr(), bytes.len(), charwise as i32, &mut h, &mut used) })?;
    let rc = unsafe { ffi::dach_dev_upload(h, device, &mut dev) };
    unsafe { ffi::dach_pma_free(h) };
    check(rc)?;
    Ok(Device { dev })
}

fn scan<P: AsRef<[u8]>>(dev: &Device, mode: i32, haystacks: &[P]) -> Result<BatchMatches, GpuError> {
    let mut text = Vec::new();
    let mut offs = Vec::with_capacity(haystacks.len() + 1);
    offs.push(0u64);
    for h in haystacks {
        text.extend_from_slice(h.as_ref());
        offs.push(text.len() as u64);
    }
    let mut out_offs = vec![0u64; haystacks.len() + 1];
    let mut cap = (text.len() / 8).max(1024) as u64;
    loop {
        let mut out = vec![ffi::DachMatch::default(); cap as usize];
        let mut needed = 0u64;
        let rc = unsafe {
            ffi::dach_scan_batch_host(dev.dev, mode, text.as_ptr(), offs.as_ptr(), haystacks.len() as u64,
                                      out.as_mut_ptr(), cap, out_offs.as_mut_ptr(), &mut needed)
        };
        match rc {
            ffi::DACH_OK => {
                out.truncate(needed as usize);
                // Match { length, end, value } (src/lib.rs:287-320): length = end - start
                let matches = out
                    .iter()
                    .map(|t| Match { length: (t.end - t.start) as usize, end: t.end as usize, value: t.value })
                    .collect();
                return Ok(BatchMatches { matches, offsets: out_offs });
            }
            ffi::DACH_OUTPUT_OVERFLOW => cap = needed, // eager API: retry with the exact size
            ffi::DACH_MATCH_KIND_MISMATCH => panic!("Error: match_kind mismatch"), // src/bytewise.rs:194-197
            _ => return Err(GpuError { code: rc, message: last_error() }),
        }
    }
}

/// A `DoubleArrayAhoCorasick<u32>` whose scans run on a B200.
pub struct GpuDoubleArrayAhoCorasick {
    dev: Device,
    match_kind: MatchKind,
}

impl GpuDoubleArrayAhoCorasick {
    /// Build with the crate (`DoubleArrayAhoCorasick::new`, builders: unchanged), hand the serialized
    /// automaton over (src/bytewise.rs:801), upload once.
    pub fn from_pma(pma: &DoubleArrayAhoCorasick<u32>, device: i32) -> Result<Self, GpuError> {
        Ok(Self { dev: upload(&pma.serialize(), false, device)?, match_kind: pma.match_kind() })
    }
    /// Batch form of `find_iter` (src/bytewise.rs:190).
    pub fn find_batch<P: AsRef<[u8]>>(&self, h: &[P]) -> Result<BatchMatches, GpuError> {
        assert!(self.match_kind.is_standard(), "Error: match_kind must be standard.");
        scan(&self.dev, ffi::DACH_FIND, h)
    }
    /// Batch form of `find_overlapping_iter` (src/bytewise.rs:292).
    pub fn find_overlapping_batch<P: AsRef<[u8]>>(&self, h: &[P]) -> Result<BatchMatches, GpuError> {
        assert!(self.match_kind.is_standard(), "Error: match_kind must be standard.");
        scan(&self.dev, ffi::DACH_FIND_OVERLAPPING, h)
    }
    /// Batch form of `find_overlapping_no_suffix_iter` (src/bytewise.rs:410).
    pub fn find_overlapping_no_suffix_batch<P: AsRef<[u8]>>(&self, h: &[P]) -> Result<BatchMatches, GpuError> {
        assert!(self.match_kind.is_standard(), "Error: match_kind must be standard.");
        scan(&self.dev, ffi::DACH_FIND_OVERLAPPING_NO_SUFFIX, h)
    }
    /// Batch form of `leftmost_find_iter` (src/bytewise.rs:547).
    pub fn leftmost_find_batch<P: AsRef<[u8]>>(&self, h: &[P]) -> Result<BatchMatches, GpuError> {
        assert!(self.match_kind.is_leftmost(), "Error: match_kind must be leftmost.");
        scan(&self.dev, ffi::DACH_LEFTMOST_FIND, h)
    }
    /// Drop-in for the lazy iterator on one haystack: scan eagerly, iterate the result.
    pub fn find_overlapping_iter<P: AsRef<[u8]>>(&self, haystack: P) -> impl Iterator<Item = Match<u32>> {
        self.find_overlapping_batch(&[haystack]).expect("GPU scan failed").matches.into_iter()
    }
}

/// A `CharwiseDoubleArrayAhoCorasick<u32>` whose scans run on a B200 (haystacks are `&str`: valid UTF-8).
pub struct GpuCharwiseDoubleArrayAhoCorasick {
    dev: Device,
    match_kind: MatchKind,
}

impl GpuCharwiseDoubleArrayAhoCorasick {
    pub fn from_pma(pma: &CharwiseDoubleArrayAhoCorasick<u32>, device: i32) -> Result<Self, GpuError> {
        Ok(Self { dev: upload(&pma.serialize(), true, device)?, match_kind: pma.match_kind() })
    }
    pub fn find_batch(&self, h: &[&str]) -> Result<BatchMatches, GpuError> {
        assert!(self.match_kind.is_standard(), "Error: match_kind must be standard.");
        scan(&self.dev, ffi::DACH_FIND, h)
    }
    pub fn find_overlapping_batch(&self, h: &[&str]) -> Result<BatchMatches, GpuError> {
        assert!(self.match_kind.is_standard(), "Error: match_kind must be standard.");
        scan(&self.dev, ffi::DACH_FIND_OVERLAPPING, h)
    }
    pub fn find_overlapping_no_suffix_batch(&self, h: &[&str]) -> Result<BatchMatches, GpuError> {
        assert!(self.match_kind.is_standard(), "Error: match_kind must be standard.");
        scan(&self.dev, ffi::DACH_FIND_OVERLAPPING_NO_SUFFIX, h)
    }
    pub fn leftmost_find_batch(&self, h: &[&str]) -> Result<BatchMatches, GpuError> {
        assert!(self.match_kind.is_leftmost(), "Error: match_kind must be leftmost.");
        scan(&self.dev, ffi::DACH_LEFTMOST_FIND, h)
    }
}

/// Many streams scanned chunk after chunk: the batch form of `find_overlapping_stepper()`
/// (src/bytewise.rs:660-729).  `state[i]` / `pos[i]` are stream i's `state_id` / `pos` and live in device
/// memory (allocated by the caller's CUDA binding); each call consumes one chunk per stream.
pub struct GpuStreamScanner<'a> {
    pub pma: &'a GpuDoubleArrayAhoCorasick,
}

impl GpuStreamScanner<'_> {
    /// # Safety
    /// All pointers are device pointers of the sizes `dach_dev_scan_stream` documents.
    pub unsafe fn consume_chunks(&self, d_text: *const u8, d_offs: *const u64, n: u64, text_bytes: u64,
                                 d_state: *mut u32, d_pos: *const u32, d_out: *mut ffi::DachMatch, out_cap: u64,
                                 d_out_offs: *mut u64, stream: *mut core::ffi::c_void) -> Result<u64, GpuError> {
        let mut needed = 0u64;
        check(ffi::dach_dev_scan_stream(self.pma.dev.dev, ffi::DACH_FIND_OVERLAPPING, d_text, d_offs, n, text_bytes,
                                        d_state, d_pos, d_out, out_cap, d_out_offs, &mut needed, stream))?;
        Ok(needed)
    }
}
