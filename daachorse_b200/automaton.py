"""Host-side mirror of the crate's public interface for the scan path.

Names, argument meaning and error behaviour follow daac-tools/daachorse 4.0.0
(file:line relative to the crate root):

    DoubleArrayAhoCorasick            src/bytewise.rs:54-68   (new :103, with_values :145,
        find_iter :190, find_overlapping_iter :292, find_overlapping_no_suffix_iter :410,
        leftmost_find_iter :547, match_kind :747, heap_bytes :764, num_states :785,
        serialize :801, deserialize :868)
    DoubleArrayAhoCorasickBuilder     src/bytewise/builder.rs:21-244
    CharwiseDoubleArrayAhoCorasick    src/charwise.rs:59-65 (same surface + num_elements :796)
    CharwiseDoubleArrayAhoCorasickBuilder  src/charwise/builder.rs
    MatchKind                         src/lib.rs:324-346
    Match                             src/lib.rs:287-320

Construction runs on the host inside libdaachorse_b200.so; every scan runs on the GPU
through the C ABI (include/daachorse_b200.h).  There is no CPU scan path: without a CUDA
device the scan methods raise DaachorseError(CUDA_ERROR).

The crate's iterators are lazy; here an iterator call scans the haystack eagerly on the
device and then yields the same Match sequence.  The ``*_batch`` methods are the
throughput interface (many haystacks per call).
"""
import ctypes as C
import enum

import numpy as np

from . import _lib

MATCH_DTYPE = np.dtype([("start", "<u4"), ("end", "<u4"), ("value", "<u4")])

FIND, FIND_OVERLAPPING, FIND_OVERLAPPING_NO_SUFFIX, LEFTMOST_FIND = range(4)


class MatchKind(enum.IntEnum):
    """src/lib.rs:324-346"""
    Standard = 0
    LeftmostLongest = 1
    LeftmostFirst = 2


class DaachorseError(Exception):
    """Mirror of errors::DaachorseError (src/errors.rs:10-22); ``code`` is the dach_status."""

    NAMES = {1: "InvalidArgument", 2: "AutomatonScale", 3: "InvalidConversion", 4: "InvalidAutomaton",
             5: "MatchKindMismatch", 6: "OutputOverflow", 7: "CudaError"}

    def __init__(self, code, msg=""):
        super().__init__("%s: %s" % (self.NAMES.get(code, code), msg))
        self.code = code


def _check(rc):
    if rc == _lib.OK:
        return
    msg = _lib.last_error()
    if rc == _lib.MATCH_KIND_MISMATCH:
        # the crate panics: assert!(self.match_kind.is_standard(), ...) (src/bytewise.rs:194-197)
        raise AssertionError(msg)
    raise DaachorseError(rc, msg)


class Match:
    """src/lib.rs:287-320: start() / end() / value()."""
    __slots__ = ("_s", "_e", "_v")

    def __init__(self, start, end, value):
        self._s, self._e, self._v = int(start), int(end), int(value)

    def start(self):
        return self._s

    def end(self):
        return self._e

    def value(self):
        return self._v

    def __eq__(self, o):
        return isinstance(o, Match) and (self._s, self._e, self._v) == (o._s, o._e, o._v)

    def __hash__(self):
        return hash((self._s, self._e, self._v))

    def __repr__(self):
        return "Match { start: %d, end: %d, value: %d }" % (self._s, self._e, self._v)


class BatchResult:
    """Matches of a batch: ``matches`` (structured array start/end/value, or an (n,3) torch
    tensor for device-resident scans) and ``offsets`` (n+1) delimiting each haystack's run."""

    def __init__(self, matches, offsets):
        self.matches = matches
        self.offsets = offsets

    def __len__(self):
        return len(self.offsets) - 1

    def triples(self, i):
        lo, hi = int(self.offsets[i]), int(self.offsets[i + 1])
        m = self.matches[lo:hi]
        if isinstance(m, np.ndarray):
            return [(int(a), int(b), int(c)) for a, b, c in zip(m["start"], m["end"], m["value"])]
        return [tuple(int(x) for x in row) for row in m.cpu().tolist()]


def _pack(items, as_str):
    bs = []
    for p in items:
        if isinstance(p, str):
            bs.append(p.encode("utf-8"))
        else:
            if as_str:
                # the charwise crate API takes &str: reject invalid UTF-8 early
                bytes(p).decode("utf-8")
            bs.append(bytes(p))
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    blob = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, dtype=np.uint8)
    return blob, offs


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a.size else None


class _Automaton:
    """Shared implementation of the two automaton classes."""

    _charwise = False

    def __init__(self, handle):
        self._h = handle
        self._devs = {}

    # -- construction ----------------------------------------------------------------------
    @classmethod
    def _build(cls, patterns, values, match_kind, num_free_blocks):
        L = _lib.load()
        blob, offs = _pack(list(patterns), cls._charwise)
        n = len(offs) - 1
        vals = None
        if values is not None:
            for v in values:
                if not (0 <= int(v) <= 0xFFFFFFFF):
                    # V::try_from(i) failed (src/bytewise/builder.rs:160-165)
                    raise DaachorseError(_lib.INVALID_CONVERSION, "value does not fit u32")
            vals = np.ascontiguousarray(values, dtype=np.uint32)
            if vals.size != n:
                raise DaachorseError(_lib.INVALID_ARGUMENT, "one value per pattern expected")
        h = C.c_void_p()
        f = L.dach_charwise_build if cls._charwise else L.dach_bytewise_build
        _check(f(_ptr(blob), C.c_void_p(offs.ctypes.data), None if vals is None else _ptr(vals), n,
                 int(match_kind), int(num_free_blocks), C.byref(h)))
        return cls(h)

    @classmethod
    def new(cls, patterns):
        """``new(patterns)``: value i is associated with patterns[i]."""
        return cls._build(patterns, None, MatchKind.Standard, 16)

    @classmethod
    def with_values(cls, patvals):
        """``with_values([(pattern, value), ...])``"""
        patvals = list(patvals)
        return cls._build([p for p, _ in patvals], [v for _, v in patvals], MatchKind.Standard, 16)

    @classmethod
    def deserialize(cls, source):
        """Returns (automaton, remaining bytes) like the crate's ``deserialize``."""
        L = _lib.load()
        buf = np.frombuffer(bytes(source), dtype=np.uint8)
        h = C.c_void_p()
        used = C.c_size_t()
        _check(L.dach_pma_deserialize(_ptr(buf), buf.size, int(cls._charwise), C.byref(h), C.byref(used)))
        return cls(h), bytes(source)[used.value:]

    # deserialize_unchecked (src/bytewise.rs:1009) maps to the checked loader on purpose
    deserialize_unchecked = deserialize

    def __del__(self):
        try:
            L = _lib.load()
            for d in self._devs.values():
                L.dach_dev_free(d)
            self._devs = {}
            if self._h:
                L.dach_pma_free(self._h)
                self._h = None
        except Exception:
            pass

    # -- introspection ------------------------------------------------------------------------
    def match_kind(self):
        return MatchKind(_lib.load().dach_pma_match_kind(self._h))

    def heap_bytes(self):
        return _lib.load().dach_pma_heap_bytes(self._h)

    def num_states(self):
        return _lib.load().dach_pma_num_states(self._h)

    def num_elements(self):
        return _lib.load().dach_pma_num_elements(self._h)

    def max_pattern_len(self):
        return _lib.load().dach_pma_max_pattern_len(self._h)

    def serialize(self):
        L = _lib.load()
        n = L.dach_pma_serialized_bytes(self._h)
        buf = np.zeros(max(n, 1), dtype=np.uint8)
        w = C.c_size_t()
        _check(L.dach_pma_serialize(self._h, C.c_void_p(buf.ctypes.data), n, C.byref(w)))
        return buf[:n].tobytes()

    # -- device ---------------------------------------------------------------------------------
    def device_handle(self, device=None):
        """Uploads the scan image to ``device`` once (default: the current CUDA device)."""
        L = _lib.load()
        if device is None:
            device = _current_device()
        d = self._devs.get(device)
        if d is None:
            d = C.c_void_p()
            _check(L.dach_dev_upload(self._h, int(device), C.byref(d)))
            self._devs[device] = d
        return d

    def set_option(self, name, value, device=None):
        _check(_lib.load().dach_dev_set_option(self.device_handle(device), name.encode(), int(value)))

    def stats(self, device=None):
        L = _lib.load()
        d = self.device_handle(device)
        return {"launches": L.dach_dev_kernel_launches(d), "scan_kernel_ms": L.dach_dev_last_scan_kernel_ms(d),
                "total_ms": L.dach_dev_last_total_ms(d), "h2d_bytes": L.dach_dev_last_h2d_bytes(d),
                "d2h_bytes": L.dach_dev_last_d2h_bytes(d), "image_bytes": L.dach_dev_image_bytes(d)}

    def _assert_mode(self, mode):
        lm = self.match_kind() != MatchKind.Standard
        if (mode == LEFTMOST_FIND) != lm:
            raise AssertionError("Error: match_kind must be %s." % ("standard" if lm else "leftmost"))

    # -- batch scans ----------------------------------------------------------------------------
    def scan_batch_host(self, mode, text, offs, out_cap=None, device=None, out=None, out_offs=None):
        """Host buffers in, host buffers out (numpy).  ``text`` uint8, ``offs`` uint64 (n+1).
        ``out`` (MATCH_DTYPE) / ``out_offs`` (uint64, n+1) may be preallocated -- e.g. views of pinned
        memory -- to keep allocation and page faults out of the call."""
        self._assert_mode(mode)
        L = _lib.load()
        d = self.device_handle(device)
        text = np.ascontiguousarray(text, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        if n < 0 or (n > 0 and int(offs.max()) > text.size):
            raise DaachorseError(_lib.INVALID_ARGUMENT, "offsets must hold n + 1 entries and stay inside the text")
        if out is not None and (out.dtype != MATCH_DTYPE or not out.flags.c_contiguous):
            raise DaachorseError(_lib.INVALID_ARGUMENT, "out must be a contiguous MATCH_DTYPE array")
        if out_offs is not None and (out_offs.dtype != np.uint64 or not out_offs.flags.c_contiguous or len(out_offs) < n + 1):
            raise DaachorseError(_lib.INVALID_ARGUMENT, "out_offs must be a contiguous uint64 array of n + 1 entries")
        if out is not None:
            cap = len(out)
        else:
            cap = int(out_cap) if out_cap else max(1024, int(text.size // 8))
        if out_offs is None:
            out_offs = np.empty(n + 1, dtype=np.uint64)
        while True:
            if out is None or len(out) < cap:
                out = np.empty(cap, dtype=MATCH_DTYPE)
            need = C.c_uint64()
            rc = L.dach_scan_batch_host(d, mode, _ptr(text), C.c_void_p(offs.ctypes.data), n,
                                        C.c_void_p(out.ctypes.data), len(out), C.c_void_p(out_offs.ctypes.data),
                                        C.byref(need))
            if rc == _lib.OUTPUT_OVERFLOW:
                if int(need.value) <= len(out):
                    raise DaachorseError(rc, "overflow reported although capacity %d >= needed %d" % (len(out), need.value))
                cap = int(need.value)
                out = None
                continue
            _check(rc)
            return BatchResult(out[: need.value], out_offs[: n + 1])

    def scan_batch_device(self, mode, text, offs, out=None, out_offs=None, stream=None):
        """Device-resident scan.  ``text`` (uint8) and ``offs`` (int64/uint64, n+1) are CUDA torch
        tensors; returns BatchResult with an (total, 3) int32-typed view of u32 triples and an
        int64 offsets tensor, both on the device.  ``out``: optional preallocated (cap, 3) int32."""
        import torch

        self._assert_mode(mode)
        L = _lib.load()
        dev = text.device.index if text.device.index is not None else torch.cuda.current_device()
        _check_device_batch(text, offs, dev, out, out_offs)
        d = self.device_handle(dev)
        n = offs.numel() - 1
        if out_offs is None:
            out_offs = torch.empty(n + 1, dtype=torch.int64, device=text.device)
        cap = out.shape[0] if out is not None else max(1024, int(text.numel() // 8))
        st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(text.device).cuda_stream)
        while True:
            if out is None or out.shape[0] < cap:
                out = torch.empty((cap, 3), dtype=torch.int32, device=text.device)
            need = C.c_uint64()
            rc = L.dach_dev_scan_batch(d, mode, C.c_void_p(text.data_ptr()), C.c_void_p(offs.data_ptr()), n,
                                       text.numel(), C.c_void_p(out.data_ptr()), out.shape[0],
                                       C.c_void_p(out_offs.data_ptr()), C.byref(need), st)
            if rc == _lib.OUTPUT_OVERFLOW:
                if int(need.value) <= cap:
                    raise DaachorseError(rc, "overflow reported although capacity %d >= needed %d" % (cap, need.value))
                cap = int(need.value)
                out = None
                continue
            _check(rc)
            return BatchResult(out[: need.value], out_offs)

    def scan_stream_device(self, mode, text, offs, state, pos=None, out=None, out_offs=None, stream=None):
        """Chunks of streams -- the batch form of ``find_stepper()`` / ``find_overlapping_stepper()``
        (src/bytewise.rs:627-729): haystack i is the next chunk of stream i.  ``state`` (CUDA int32/uint32,
        n entries) holds each stream's state id and is updated in place; ``pos`` (optional, n entries) is
        the stream position of each chunk's first byte and is added to the reported positions.  For every
        byte: consume(byte), then matches().  Returns a BatchResult of device tensors."""
        import torch

        self._assert_mode(mode)
        L = _lib.load()
        dev = text.device.index if text.device.index is not None else torch.cuda.current_device()
        _check_device_batch(text, offs, dev, out, out_offs, state, pos)
        d = self.device_handle(dev)
        n = offs.numel() - 1
        if out_offs is None:
            out_offs = torch.empty(n + 1, dtype=torch.int64, device=text.device)
        cap = out.shape[0] if out is not None else max(1024, int(text.numel() // 8))
        st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(text.device).cuda_stream)
        state_in = state.clone()  # the call advances `state` even when the output overflows
        while True:
            if out is None or out.shape[0] < cap:
                out = torch.empty((cap, 3), dtype=torch.int32, device=text.device)
            need = C.c_uint64()
            rc = L.dach_dev_scan_stream(d, mode, C.c_void_p(text.data_ptr()), C.c_void_p(offs.data_ptr()), n, text.numel(),
                                        C.c_void_p(state.data_ptr()), C.c_void_p(pos.data_ptr()) if pos is not None else None,
                                        C.c_void_p(out.data_ptr()), out.shape[0], C.c_void_p(out_offs.data_ptr()),
                                        C.byref(need), st)
            if rc == _lib.OUTPUT_OVERFLOW:
                if int(need.value) <= cap:
                    raise DaachorseError(rc, "overflow reported although capacity %d >= needed %d" % (cap, need.value))
                cap = int(need.value)
                out = None
                state.copy_(state_in)
                continue
            _check(rc)
            return BatchResult(out[: need.value], out_offs)

    def job(self, device=None):
        """An asynchronous scan with its own workspace (dach_job_*): ``scan`` and ``place`` only enqueue work,
        ``wait`` blocks.  Several jobs of one automaton overlap across streams and host threads."""
        return Job(self, device)

    def _batch(self, mode, haystacks):
        blob, offs = _pack(list(haystacks), self._charwise)
        return self.scan_batch_host(mode, blob, offs)

    def find_batch(self, haystacks):
        return self._batch(FIND, haystacks)

    def find_overlapping_batch(self, haystacks):
        return self._batch(FIND_OVERLAPPING, haystacks)

    def find_overlapping_no_suffix_batch(self, haystacks):
        return self._batch(FIND_OVERLAPPING_NO_SUFFIX, haystacks)

    def leftmost_find_batch(self, haystacks):
        return self._batch(LEFTMOST_FIND, haystacks)

    # -- the crate's iterator surface ---------------------------------------------------------------
    def _iter(self, mode, haystack):
        self._assert_mode(mode)  # the crate asserts when the iterator is created
        r = self._batch(mode, [haystack])
        m = r.matches
        return iter([Match(a, b, c) for a, b, c in zip(m["start"], m["end"], m["value"])])

    def find_iter(self, haystack):
        return self._iter(FIND, haystack)

    def find_overlapping_iter(self, haystack):
        return self._iter(FIND_OVERLAPPING, haystack)

    def find_overlapping_no_suffix_iter(self, haystack):
        return self._iter(FIND_OVERLAPPING_NO_SUFFIX, haystack)

    def leftmost_find_iter(self, haystack):
        return self._iter(LEFTMOST_FIND, haystack)


class Job:
    """dach_job: one in-flight scan of device-resident buffers (include/daachorse_b200.h, "asynchronous scans")."""

    def __init__(self, pma, device=None):
        import torch

        self._pma = pma
        self._dev = torch.cuda.current_device() if device is None else int(device)
        self._h = C.c_void_p()
        _check(_lib.load().dach_job_create(pma.device_handle(self._dev), C.byref(self._h)))
        self._keep = None

    def __del__(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().dach_job_free(self._h)
            self._h = C.c_void_p()

    @staticmethod
    def _stream(stream, device):
        import torch

        if stream is None:
            return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        return C.c_void_p(stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))

    def scan(self, mode, text, offs, cap_matches, stream=None):
        """Enqueue the scan of ``text`` (uint8 CUDA tensor) / ``offs`` (int64, n+1) on ``stream``."""
        self._pma._assert_mode(mode)
        _check_device_batch(text, offs, self._dev)
        self._keep = (text, offs)  # the kernels read them after this call returns
        _check(_lib.load().dach_job_scan(self._h, mode, C.c_void_p(text.data_ptr()), C.c_void_p(offs.data_ptr()),
                                         offs.numel() - 1, text.numel(), int(cap_matches), self._stream(stream, text.device)))

    def place(self, out, out_offs, base=None, stream=None):
        """Enqueue the gather into ``out`` ((cap, 3) int32) / ``out_offs`` (int64, n+1); ``base``: optional
        1-element int64 CUDA tensor, index of the first match in ``out``."""
        _check_device_batch(self._keep[0], self._keep[1], out.device.index, out, out_offs)
        if base is not None and (not base.is_cuda or base.dtype not in (torch_int64(),) or base.numel() < 1):
            raise DaachorseError(_lib.INVALID_ARGUMENT, "base must be a 1-element int64 CUDA tensor")
        self._out = (out, out_offs, base)
        _check(_lib.load().dach_job_place(self._h, C.c_void_p(out.data_ptr()), out.shape[0], C.c_void_p(out_offs.data_ptr()),
                                          C.c_void_p(base.data_ptr()) if base is not None else None,
                                          self._stream(stream, out.device)))

    def wait(self):
        """Block until the placement is done; returns the number of matches (raises on overflow)."""
        need = C.c_uint64()
        _check(_lib.load().dach_job_wait(self._h, C.byref(need)))
        return int(need.value)

    def scan_kernel_ms(self):
        return _lib.load().dach_job_scan_kernel_ms(self._h)

    def push_ms(self):
        return _lib.load().dach_job_push_ms(self._h)

    def times(self):
        """ms since the automaton's first job scan on this device: (scan start, scan end, push start, push end)."""
        buf = (C.c_double * 4)()
        _check(_lib.load().dach_job_times(self._h, C.byref(buf)))
        return tuple(float(x) for x in buf)


def torch_int64():
    import torch

    return torch.int64


def _check_device_batch(text, offs, dev_index, out=None, out_offs=None, state=None, pos=None):
    """Raw pointers cross the C ABI: a wrong dtype, stride or device would be silent garbage or a device fault."""
    import torch

    def bad(msg):
        raise DaachorseError(_lib.INVALID_ARGUMENT, msg)

    n = offs.numel() - 1
    if n < 0:
        bad("offs must hold n + 1 entries")
    for name, t, dtypes in (("text", text, (torch.uint8,)), ("offs", offs, (torch.int64, torch.uint64)),
                            ("out", out, (torch.int32, torch.uint32)), ("out_offs", out_offs, (torch.int64, torch.uint64)),
                            ("state", state, (torch.int32, torch.uint32)), ("pos", pos, (torch.int32, torch.uint32))):
        if t is None:
            continue
        if not t.is_cuda or t.device.index != dev_index:
            bad("%s must be a CUDA tensor on cuda:%d" % (name, dev_index))
        if t.dtype not in dtypes:
            bad("%s has dtype %s, expected one of %s" % (name, t.dtype, dtypes))
        if not t.is_contiguous():
            bad("%s must be contiguous" % name)
    if out is not None and (out.dim() != 2 or out.shape[1] != 3):
        bad("out must have shape (capacity, 3)")
    if out_offs is not None and out_offs.numel() < n + 1:
        bad("out_offs must hold n + 1 entries")
    for name, t in (("state", state), ("pos", pos)):
        if t is not None and t.numel() != n:
            bad("%s must hold n entries" % name)


def _current_device():
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


class DoubleArrayAhoCorasick(_Automaton):
    """Byte-wise double-array Aho-Corasick automaton (src/bytewise.rs:54-68)."""
    _charwise = False


class CharwiseDoubleArrayAhoCorasick(_Automaton):
    """Char-wise double-array Aho-Corasick automaton (src/charwise.rs:59-65)."""
    _charwise = True


class _Builder:
    _cls = None

    def __init__(self):
        self._kind = MatchKind.Standard
        self._nfb = 16  # src/bytewise/builder.rs:61

    @classmethod
    def new(cls):
        return cls()

    def match_kind(self, kind):
        self._kind = MatchKind(kind)
        return self

    def num_free_blocks(self, n):
        assert n >= 1  # src/bytewise/builder.rs:113
        self._nfb = int(n)
        return self

    def build(self, patterns):
        return self._cls._build(patterns, None, self._kind, self._nfb)

    def build_with_values(self, patvals):
        patvals = list(patvals)
        return self._cls._build([p for p, _ in patvals], [v for _, v in patvals], self._kind, self._nfb)


class DoubleArrayAhoCorasickBuilder(_Builder):
    """src/bytewise/builder.rs:21-244"""
    _cls = DoubleArrayAhoCorasick


class CharwiseDoubleArrayAhoCorasickBuilder(_Builder):
    """src/charwise/builder.rs"""
    _cls = CharwiseDoubleArrayAhoCorasick
