"""Drop-in replacements for blackbox.py's reduction steps, running on a B200.

Same names, argument meaning and error behaviour (Python exceptions) as the reference:

    gain_corr(data, header, tel=None)                               blackbox.py:7442
    os_corr(data, header, imgtype, xbin=1, ybin=1, data_limit=2000, tel=None)   :6407
    mask_init(data, header, filt, imgtype)                          blackbox.py:4375
    cosmics_corr(data, header, data_mask, header_mask)              blackbox.py:4259
    xtalk_corr(data, crosstalk_file, data_mask=None)                blackbox.py:7138
    detect_cosmics(indat, inmask=None, sigclip=..., ...)            astroscrappy 1.0.8
    master_combine(frames, imgtype, ...)      arithmetic core of master_prep, :4908-4984

``data`` may be a numpy array (copied to the GPU, results copied back / mutated in place
exactly where the reference mutates) or a CUDA ``torch.Tensor`` (nothing leaves HBM).
``header`` is any mapping; astropy ``Header`` objects get ``(value, comment)`` tuples.
The module-global ``tel`` mirrors the reference's global of that name (blackbox.py:141).

All arithmetic happens in libbbx.so (hand-written sm_100a kernels, include/bbx.h); there is
no CPU fallback.  The only host numerics are FITPACK's smoothing spline (hostfit.py).
"""
import ctypes as C
import logging
import os

import numpy as np
import torch

from . import _lib, fitsio, hostfit, set_bb
from ._lib import BbxMaskBits, call, query
from .geometry import Geometry, define_sections  # noqa: F401  (define_sections: part of the mirrored surface)
from .set_bb import get_par

tel = None          # module-global telescope name, as in the reference (blackbox.py:141)
log = logging.getLogger(__name__)

_bpm_registry = {}  # filter -> bad-pixel mask (numpy or CUDA tensor); see set_bad_pixel_mask
_bpm_cache = {}     # path -> (mtime, size, uint8 CUDA tensor) of the masks mask_init read from disk

MASK_MORPH_SPARSE = True   # mask_init: seed-list driven morphology (False: dense passes; same result)


# -------------------------------------------------------------------------------------------
# plumbing
# -------------------------------------------------------------------------------------------
def _device():
    if not torch.cuda.is_available():
        raise _lib.BbxUnavailable('no CUDA device: blackbox_b200 has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _to_dev(a, dtype=None):
    """numpy / tensor -> contiguous CUDA tensor (no copy if already there)."""
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        t = a if a.is_cuda else a.to(_device())
    else:
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint16:
            t = torch.from_numpy(a.view(np.int16)).to(_device()).view(torch.uint16)
        else:
            t = torch.from_numpy(a).to(_device())
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _harr(values, ctype):
    return (ctype * len(values))(*values)


def _set(header, key, value, comment=None):
    if header is None:
        return
    if comment is not None and hasattr(header, 'comments'):
        header[key] = (value, comment)
    else:
        header[key] = value


def _raw_type(t):
    if t.dtype == torch.uint16:
        return 0
    if t.dtype == torch.float32:
        return 1
    # int16 is what a BZERO = 0 frame decodes to: its negative counts must not be read as 32768+
    raise TypeError('raw frame must be uint16 (counts) or float32, got {}'.format(t.dtype))


def _inplace_target(a, dtype, who):
    """The tensor an in-place entry point works on: numpy arrays are copied to the device (and
    copied back by the caller); a torch tensor must already BE the thing to modify -- a CUDA,
    contiguous tensor of ``dtype`` -- because a silent copy would swallow the result."""
    if isinstance(a, torch.Tensor):
        if not a.is_cuda or not a.is_contiguous() or a.dtype != dtype:
            raise TypeError('{}: works in place and needs a contiguous CUDA {} tensor (got {} on {}, '
                            'contiguous={}); pass a numpy array or convert first'.format(
                                who, dtype, a.dtype, a.device, a.is_contiguous()))
        return a
    return _to_dev(a, dtype)


def _bits(tel_):
    return BbxMaskBits.from_dict(get_par(set_bb.mask_value, tel_))


def _tel_kind(tel_):
    return 0 if str(tel_)[0:2] == 'ML' else 1


# -------------------------------------------------------------------------------------------
# gain
# -------------------------------------------------------------------------------------------
def gain_corr(data, header, tel=None):
    """In place ``data[chan] *= gain[chan]``; header GAIN1..16 (blackbox.py:7442-7465)."""
    gain = get_par(set_bb.gain, tel)
    is_np = isinstance(data, np.ndarray)
    if is_np and data.dtype != np.float32:
        raise TypeError('gain_corr expects a float32 frame (as read_hdulist(dtype="float32"))')
    t = _inplace_target(data, torch.float32, 'gain_corr')
    geom = Geometry.from_raw_shape(tuple(t.shape), tel=tel) if _has_overscan(t.shape, tel) \
        else _reduced_geometry(tuple(t.shape), tel)
    g = geom.as_struct()
    call('bbx_gain_corr', _ptr(t), C.byref(g), _harr([float(x) for x in gain], C.c_float), _stream())
    if is_np:
        data[...] = t.cpu().numpy()
    for i in range(geom.nchans):
        _set(header, 'GAIN{}'.format(i + 1), gain[i],
             '[e-/ADU] gain applied to channel {}'.format(i + 1))


def _has_overscan(shape, tel_, xbin=1, ybin=1):
    ysc = get_par(set_bb.ysize_chan, tel_) // ybin
    xsc = get_par(set_bb.xsize_chan, tel_) // xbin
    return shape[0] > set_bb.ny * ysc and shape[1] > set_bb.nx * xsc


def _reduced_geometry(shape, tel_):
    """Geometry of a frame without overscans (tile == data section)."""
    ny, nx = get_par(set_bb.ny, tel_), get_par(set_bb.nx, tel_)
    H, W = shape
    if H % ny or W % nx:
        raise ValueError('frame {} is not divisible into {} x {} channels'.format(shape, ny, nx))
    dy, dx = H // ny, W // nx
    return Geometry(H, W, ny, nx, dy, dx, dy, dx, 0, 0, 0, (0, dy), (0, dy))


# -------------------------------------------------------------------------------------------
# overscan
# -------------------------------------------------------------------------------------------
class OverscanState:
    """Device-resident intermediates and header scalars of one frame.

    Every field is a view into ONE device buffer.  The fields the host needs for the spline
    decision (``HOST_FIELDS``, at the front of the buffer) travel in a single device-to-host copy
    into a pinned mirror right after the overscan stage (``fetch_async`` / ``host``); every
    scalar that ends up in the FITS header (``HEADER_FIELDS``, at the end of the buffer: the
    os_corr keywords, the saturation levels, NOBJ-SAT, NCOSMICS, the LACosmic and morphology
    status words, the per-bit mask counts) travels in ONE copy at the end of the chain
    (``fetch_header_async`` / ``hhost``) -- no ``.item()`` anywhere on the frame path."""

    HOST_FIELDS = ('fit_status', 'need_spline', 'hos_n', 'hos_mean', 'hos_std', 'oscan')
    HEADER_FIELDS = ('vos_coef', 'biasm', 'vfit_ok', 'std_vos', 'satlevel', 'means', 'mstatus', 'nobj',
                     'ncosmic', 'lacinfo', 'mcounts')

    def __init__(self, geom, device, niter=4):
        n, dy, nc = geom.nchans, geom.dy, geom.xsize_chan
        f64, f32, i32, i64, u8 = torch.float64, torch.float32, torch.int32, torch.int64, torch.uint8
        fields = [('fit_status', (n,), i32), ('need_spline', (n, nc), u8), ('hos_n', (n, nc), i32),
                  ('hos_mean', (n, nc), f32), ('hos_std', (n, nc), f32), ('oscan', (n, nc), f64),
                  ('mean_vos', (n, dy), f64), ('vos_fit', (n, dy), f64), ('satcnt', (n, 2, nc), i32),
                  ('dlevel', (n,), f64), ('satcol', (n, nc), u8),
                  # ---- header block (HEADER_FIELDS, contiguous) ----
                  ('vos_coef', (n, 8), f64), ('biasm', (n,), f64), ('vfit_ok', (n,), i32),
                  ('std_vos', (n,), f64), ('satlevel', (n,), f64),
                  ('means', (2,), f64),                       # BIASMEAN, RDNOISE
                  ('mstatus', (2,), i32),                     # mask morphology status (bbx_mask_morph_sparse)
                  ('nobj', (1,), i32), ('ncosmic', (1,), i32),
                  ('lacinfo', (4 + max(int(niter), 1),), i64),   # bbx_lacosmic out_info
                  ('mcounts', (136,), i64)]                   # [0:8] pixels per mask bit (mask_header); rest: scratch of bbx_xtalk_counts
        self.geom = geom
        self._layout = {}
        off = 0
        for name, shape, dt in fields:
            if name == self.HEADER_FIELDS[0]:
                self._hdr_off = off
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            self._layout[name] = (off, nbytes, shape, dt)
            off = (off + nbytes + 15) // 16 * 16
            if name == self.HOST_FIELDS[-1]:
                self._host_bytes = off
        self._hdr_bytes = off - self._hdr_off
        self.buf = torch.zeros(off, dtype=u8, device=device)
        for name, (o, nb, shape, dt) in self._layout.items():
            setattr(self, name, self.buf[o:o + nb].view(dt).view(shape))
        self._pinned = None
        self._pinned_hdr = None

    def fetch_async(self):
        """Enqueue the copy of the host-side fields into the pinned mirror (current stream)."""
        if self._pinned is None:
            self._pinned = torch.empty(self._host_bytes, dtype=torch.uint8).pin_memory()
        self._pinned.copy_(self.buf[:self._host_bytes], non_blocking=True)

    def host(self, name):
        """numpy view of a HOST_FIELDS member in the pinned mirror (valid once the stream that
        ran ``fetch_async`` has been synchronised)."""
        o, nb, shape, dt = self._layout[name]
        return self._pinned[o:o + nb].view(dt).view(shape).numpy()

    def fetch_header_async(self):
        """Enqueue the copy of the header block into its pinned mirror (current stream)."""
        if self._pinned_hdr is None:
            self._pinned_hdr = torch.empty(self._hdr_bytes, dtype=torch.uint8).pin_memory()
        self._pinned_hdr.copy_(self.buf[self._hdr_off:self._hdr_off + self._hdr_bytes], non_blocking=True)

    def hhost(self, name):
        """numpy view of a HEADER_FIELDS member in the pinned header mirror (valid once the stream
        that ran ``fetch_header_async`` has been synchronised)."""
        o, nb, shape, dt = self._layout[name]
        o -= self._hdr_off
        return self._pinned_hdr[o:o + nb].view(dt).view(shape).numpy()


def overscan_enqueue(raw_t, geom, tel_, gain=None, data_limit=2000, state=None):
    """Enqueue every overscan statistics / fit kernel for one raw frame on the current
    stream (no synchronisation).  ``gain``: list of 16 gains to apply on the fly (raw counts)
    or None if ``raw_t`` is already gain-corrected.  Returns the OverscanState."""
    st = state if state is not None else OverscanState(geom, raw_t.device)
    g = geom.as_struct()
    rt = _raw_type(raw_t)
    gain_h = _harr([float(x) for x in gain], C.c_float) if gain is not None else None
    s = _stream()
    kind = _tel_kind(tel_)
    call('bbx_vos_rowstats', _ptr(raw_t), rt, C.byref(g), gain_h, 3.0, 5, _ptr(st.mean_vos), s)
    call('bbx_vos_fit', _ptr(st.mean_vos), C.byref(g), int(get_par(set_bb.voscan_poldeg, tel_)), 5.0,
         _ptr(st.vos_fit), _ptr(st.vos_coef), _ptr(st.biasm), _ptr(st.vfit_ok), s)
    sat_e = (np.array(get_par(set_bb.satlevel, tel_), dtype=np.float64) *
             np.array(get_par(set_bb.gain, tel_), dtype=np.float64))
    sat_e_h = _harr([float(x) for x in sat_e], C.c_double)
    if kind == 1:
        lim = get_par(set_bb.hos_sat_ypix_lim, tel_)
        call('bbx_hos_satcount', _ptr(raw_t), rt, C.byref(g), gain_h, _ptr(st.vos_fit), sat_e_h,
             int(lim[0]), int(lim[1]), _ptr(st.satcnt), s)
    call('bbx_hos_stats', _ptr(raw_t), rt, C.byref(g), gain_h, _ptr(st.vos_fit), kind,
         float(data_limit), _ptr(st.satcnt), _ptr(st.dlevel), _ptr(st.hos_mean), _ptr(st.hos_std),
         _ptr(st.hos_n), _ptr(st.satcol), s)
    call('bbx_vos_std', _ptr(raw_t), rt, C.byref(g), gain_h, _ptr(st.vos_fit), _ptr(st.dlevel),
         _ptr(st.std_vos), s)
    split_chan = 8 if tel_ == 'BG2' else -1
    call('bbx_hos_fit', _ptr(st.hos_mean), _ptr(st.hos_std), _ptr(st.hos_n), _ptr(st.satcol),
         C.byref(g), kind, split_chan, 654, _ptr(st.oscan), _ptr(st.need_spline),
         _ptr(st.fit_status), s)
    call('bbx_satlevels', sat_e_h, _ptr(st.biasm), _ptr(st.satlevel), s)
    return st


def overscan_resolve_spline(st, strict, fetched=False):
    """Host step: evaluate FITPACK's smoothing spline for the columns that need it (all
    channels if ``strict``) and patch ``st.oscan``.  Unless ``fetched`` (the caller has run
    ``st.fetch_async()`` and synchronised) this synchronises the current stream.  Returns the
    number of patched columns.  Raises RuntimeError if a polynomial fit had too few points
    (the reference raises from np.polyfit in that case)."""
    if not fetched:
        st.fetch_async()
        torch.cuda.current_stream().synchronize()
    status = st.host('fit_status')
    if status.any():
        raise RuntimeError('horizontal-overscan polynomial fit failed for channel(s) {} '
                           '(too few valid columns)'.format(
                               [int(i) + 1 for i in np.nonzero(status)[0]]))
    need = st.host('need_spline').astype(bool)
    chans = range(st.geom.nchans) if strict else np.nonzero(need.any(axis=1))[0]
    if len(chans) == 0:
        return 0
    mean, std, n, oscan = st.host('hos_mean'), st.host('hos_std'), st.host('hos_n'), st.host('oscan')
    o, _, _, _ = st._layout['oscan']
    patched = 0
    for i in chans:
        cols = np.nonzero(need[i])[0]
        spl = hostfit.hos_spline(mean[i], std[i], n[i])
        if len(cols):
            oscan[i, cols] = spl[cols]
            # the patched row goes back from the pinned mirror (one small async copy per channel)
            st.oscan[i].copy_(st._pinned[o:o + oscan.nbytes].view(torch.float64).view(oscan.shape)[i],
                              non_blocking=True)
            patched += len(cols)
    return patched


def apply_enqueue(raw_t, geom, tel_, st=None, gain=None, mbias=None, mflat=None, bpm=None,
                  want_mask=False, out_img=None, out_mask=None, mwork=None):
    """Enqueue the fused per-pixel pass (include/bbx.h: bbx_reduce_apply).  With ``mwork`` (a
    MaskWork) the saturated pixels are also appended to its seed list for the sparse mask
    morphology."""
    g = geom.as_struct()
    RH, RW = geom.red_shape
    if out_img is None:
        out_img = torch.empty((RH, RW), dtype=torch.float32, device=raw_t.device)
    if want_mask and out_mask is None:
        out_mask = torch.empty((RH, RW), dtype=torch.uint8, device=raw_t.device)
    bits = _bits(tel_)
    gain_h = _harr([float(x) for x in gain], C.c_float) if gain is not None else None
    call('bbx_reduce_apply', _ptr(raw_t), _raw_type(raw_t), C.byref(g), gain_h,
         _ptr(st.vos_fit) if st is not None else None, _ptr(st.oscan) if st is not None else None,
         _ptr(mbias), _ptr(mflat), _ptr(bpm),
         _ptr(st.satlevel) if (st is not None and want_mask) else None,
         C.byref(bits), _ptr(out_img), _ptr(out_mask) if want_mask else None,
         _ptr(mwork.seeds) if (mwork is not None and want_mask) else None,
         _ptr(mwork.seed_count) if (mwork is not None and want_mask) else None,
         int(mwork.seed_cap) if mwork is not None else 0, _stream())
    return out_img, out_mask


def fusable(geom, raw_t, *tensors):
    """bbx_reduce_apply_scan needs the 4-pixel-aligned layout: channel width, tile width and frame
    width multiples of 4, 16-byte aligned buffers (torch allocations are), at most 65535 x 32 rows."""
    if geom.xsize_chan % 4 or geom.dx % 4 or geom.W % 4 or geom.red_shape[0] > 65535 * 32:
        return False
    return all(t is None or t.data_ptr() % 16 == 0 for t in (raw_t,) + tensors)


def apply_scan_enqueue(raw_t, geom, tel_, st, gain, mbias, mflat, bpm, out_img, out_mask, mwork, crmask,
                       sigclip, sigfrac, objlim, niter, lwork, readnoise=0.0, readnoise_dev=None):
    """Enqueue the fused per-pixel pass TOGETHER with the dense Laplacian scan of detect_cosmics'
    first iteration (include/bbx.h: bbx_reduce_apply_scan).  To be followed by
    ``mask_morph_enqueue(..., track=(out_img, lwork))`` and ``lacosmic_enqueue(..., mode=LAC_FUSED)``
    with the same thresholds, work buffer and cosmic-ray mask."""
    g = geom.as_struct()
    bits = _bits(tel_)
    gain_h = _harr([float(x) for x in gain], C.c_float) if gain is not None else None
    call('bbx_reduce_apply_scan', _ptr(raw_t), _raw_type(raw_t), C.byref(g), gain_h, _ptr(st.vos_fit), _ptr(st.oscan),
         _ptr(mbias), _ptr(mflat), _ptr(bpm), _ptr(st.satlevel), C.byref(bits), _ptr(out_img), _ptr(out_mask),
         _ptr(mwork.seeds), _ptr(mwork.seed_count), int(mwork.seed_cap), _ptr(crmask),
         float(np.float32(sigclip)), float(np.float32(sigfrac)), float(np.float32(objlim)),
         float(np.float32(readnoise)), _ptr(readnoise_dev), int(niter), _ptr(lwork.buf), _ptr(lwork.info), _stream())
    return out_img, out_mask


def apply_stats_enqueue(raw_t, geom, tel_, st, gain, mbias, mflat, bpm, out_img, out_mask, mwork, niter, lwork):
    """Enqueue the fused per-pixel pass with the statistics of detect_cosmics' background level taken
    on the way (include/bbx.h: bbx_reduce_apply_stats).  To be followed by ``mask_morph_enqueue(...,
    track=(out_img, lwork))`` and ``lacosmic_enqueue(..., mode=LAC_STATS)`` with the same work buffer."""
    g = geom.as_struct()
    bits = _bits(tel_)
    gain_h = _harr([float(x) for x in gain], C.c_float) if gain is not None else None
    call('bbx_reduce_apply_stats', _ptr(raw_t), _raw_type(raw_t), C.byref(g), gain_h, _ptr(st.vos_fit), _ptr(st.oscan),
         _ptr(mbias), _ptr(mflat), _ptr(bpm), _ptr(st.satlevel), C.byref(bits), _ptr(out_img), _ptr(out_mask),
         _ptr(mwork.seeds), _ptr(mwork.seed_count), int(mwork.seed_cap), int(niter), _ptr(lwork.buf), _ptr(lwork.info),
         _stream())
    return out_img, out_mask


def os_corr(data, header, imgtype, xbin=1, ybin=1, data_limit=2000, tel=None, strict=True,
            return_state=False):
    """Overscan correction; returns the cropped float32 frame and fills the header keywords
    BIAS{i}A{n}, VFITOK{i}, BIASM{i}, RDN{i}, BIASMEAN, RDNOISE (blackbox.py:6407-6879).

    ``data``: float32 raw frame, already gain-corrected (as after gain_corr), or a uint16 raw
    frame (the gain is then applied on the fly).  Unlike the reference the input array is
    left untouched (the reference subtracts the overscan from it as a side effect that no
    caller uses).
    """
    is_np = isinstance(data, np.ndarray)
    raw_t = _to_dev(data)
    geom = Geometry.from_raw_shape(tuple(raw_t.shape), xbin=xbin, ybin=ybin, tel=tel)
    gain = get_par(set_bb.gain, tel) if _raw_type(raw_t) == 0 else None
    st = overscan_enqueue(raw_t, geom, tel, gain=gain, data_limit=data_limit)
    overscan_resolve_spline(st, strict)
    out, _ = apply_enqueue(raw_t, geom, tel, st=st, gain=gain)
    fill_os_header(header, st)
    if return_state:
        return (out.cpu().numpy() if is_np else out), st
    return out.cpu().numpy() if is_np else out


def fill_os_header(header, st, fetched=False):
    """Header keywords of os_corr from the state's header block.  Unless ``fetched`` (the caller
    has run ``st.fetch_header_async()`` and synchronised that stream) this enqueues the copy and
    synchronises the current stream."""
    if not fetched:
        st.fetch_header_async()
        torch.cuda.current_stream().synchronize()
    coef, ok = st.hhost('vos_coef'), st.hhost('vfit_ok')
    biasm, rdn = st.hhost('biasm'), st.hhost('std_vos')
    deg = int(set_bb.voscan_poldeg)
    n = st.geom.nchans
    for i in range(n):
        for nc in range(deg + 1):
            v = float(coef[i, nc])
            _set(header, 'BIAS{}A{}'.format(i + 1, nc), v if np.isfinite(v) else 'None',
                 '[e-] channel {} vert. overscan A{} polyfit coeff'.format(i + 1, nc))
        _set(header, 'VFITOK{}'.format(i + 1), bool(ok[i]),
             'channel {} vert. overscan polyfit finite?'.format(i + 1))
    for i in range(n):
        _set(header, 'BIASM{}'.format(i + 1), float(biasm[i]),
             '[e-] channel {} mean vertical overscan'.format(i + 1))
    for i in range(n):
        _set(header, 'RDN{}'.format(i + 1), float(rdn[i]),
             '[e-] channel {} sigma (STD) vertical overscan'.format(i + 1))
    _set(header, 'BIASMEAN', float(np.nanmean(biasm)), '[e-] average all channel means vert. overscan')
    _set(header, 'RDNOISE', float(np.nanmean(rdn)), '[e-] average all channel sigmas vert. overscan')


# -------------------------------------------------------------------------------------------
# bias / flat (blackbox.py:1679, 1825)
# -------------------------------------------------------------------------------------------
def subtract_mbias(data, data_mbias):
    """In place ``data -= data_mbias`` (float32)."""
    return _binary_inplace(data, data_mbias, 0)


def divide_mflat(data, data_mflat):
    """In place ``data /= data_mflat`` (float32 IEEE division)."""
    return _binary_inplace(data, data_mflat, 1)


def _binary_inplace(a, b, op):
    is_np = isinstance(a, np.ndarray)
    ta, tb = _inplace_target(a, torch.float32, 'subtract_mbias / divide_mflat'), _to_dev(b, torch.float32)
    if ta.shape != tb.shape:
        raise ValueError('shape mismatch {} vs {}'.format(tuple(ta.shape), tuple(tb.shape)))
    call('bbx_binary_inplace', _ptr(ta), _ptr(tb), ta.numel(), op, _stream())
    if is_np:
        a[...] = ta.cpu().numpy()
        return a
    return ta


# -------------------------------------------------------------------------------------------
# mask
# -------------------------------------------------------------------------------------------
def set_bad_pixel_mask(filt, bpm):
    """Register the bad-pixel mask of filter ``filt`` (the reference reads
    set_bb.bad_pixel_mask with 'bpm' -> 'bpm_<filt>' from disk, blackbox.py:4386-4398)."""
    if bpm is None:
        _bpm_registry.pop(filt, None)
    else:
        _bpm_registry[filt] = bpm


class MaskWork:
    """Scratch buffers of the mask morphology for one frame shape (reusable)."""

    def __init__(self, H, W, device, status=None, nobj=None):
        """``status`` (int32[2]) / ``nobj`` (int32[1]): device tensors to use for the status words
        and NOBJ-SAT (FramePipeline points them into the frame's header block)."""
        self.H, self.W = H, W
        self.holes = torch.empty(query('bbx_fill_holes_work_bytes', H, W), dtype=torch.uint8, device=device)
        # the hole-filling state image starts as "background everywhere"; the sparse morphology
        # leaves it like that after every frame (no memset per frame), the dense one does not
        self.holes[:H * W].fill_(1)
        self.labels = torch.empty(H * W, dtype=torch.int32, device=device)
        self.unconverged = torch.zeros(1, dtype=torch.int32, device=device)
        self.nobj = nobj if nobj is not None else torch.zeros(1, dtype=torch.int32, device=device)
        self.seed_cap = H * W // 8 + 1024
        self.seeds = torch.empty(self.seed_cap, dtype=torch.int32, device=device)
        self.seed_count = torch.zeros(1, dtype=torch.int32, device=device)
        self.status = status if status is not None else torch.zeros(2, dtype=torch.int32, device=device)


def mask_morph_enqueue(mask_t, tel_, work, count_objects=True, rounds=4096, sparse=True, track=None):
    """Enqueue crosstalk-victim / saturated-connected / NOBJ-SAT / fill_sat_holes on a mask
    that carries the saturation marker written by bbx_reduce_apply.  ``sparse``: driven by the
    seed list in ``work`` (filled by apply_enqueue(..., mwork=work)); otherwise dense passes.
    ``track``: (reduced image, LacosmicWork) after ``apply_scan_enqueue`` -- the morphology then
    keeps LACosmic's background statistics right for the pixels it masks."""
    H, W = mask_t.shape
    bits = _bits(tel_)
    s = _stream()
    if sparse:
        img_t, lwork = track if track is not None else (None, None)
        call('bbx_mask_morph_sparse_track', _ptr(mask_t), H, W, H // 2, W // 8, C.byref(bits), _ptr(work.seeds),
             _ptr(work.seed_count), int(work.seed_cap), _ptr(work.holes), _ptr(work.labels), _ptr(work.nobj),
             int(rounds), _ptr(work.status), _ptr(img_t), _ptr(lwork.buf) if lwork is not None else None, 1, s)
        return
    if track is not None:
        raise ValueError('mask_morph_enqueue: tracking needs the sparse morphology')
    work.status.zero_()
    call('bbx_mask_sat_neighbours', _ptr(mask_t), H, W, H // 2, W // 8, C.byref(bits), s)
    if count_objects:
        call('bbx_count_objects', _ptr(mask_t), 0x80, H, W, _ptr(work.labels), _ptr(work.nobj), s)
    call('bbx_fill_sat_holes', _ptr(mask_t), H, W, C.byref(bits), _ptr(work.holes), min(rounds, 64),
         _ptr(work.unconverged), s)


def mask_morph_finish(mask_t, tel_, work, rounds=1024, sparse=True):
    """Synchronise.  Sparse: returns (NOBJ-SAT, ok); ok False means the mask is not final and
    the dense path has to be run on a fresh seed mask.  Dense: continues the hole filling until
    it has converged."""
    H, W = mask_t.shape
    if sparse:
        return int(work.nobj.item()), int(work.status[0].item()) == 0
    bits = _bits(tel_)
    while int(work.unconverged.item()) != 0:
        call('bbx_fill_holes_more', _ptr(mask_t), H, W, C.byref(bits), _ptr(work.holes), rounds,
             _ptr(work.unconverged), _stream())
    work.holes[:H * W].fill_(1)              # the dense passes used the state image: as the sparse path expects it again
    return int(work.nobj.item()), True


def _load_bpm(filt, shape, device):
    """The bad-pixel mask of filter ``filt`` as mask_init finds it (blackbox.py:4386-4398):
    ``set_bb.bad_pixel_mask`` with 'bpm' -> 'bpm_<filt>', fpacked or not (``already_exists``); a
    mask registered with ``set_bad_pixel_mask`` takes precedence.  A missing file is a logged
    warning and a mask of zeros, as in the reference -- never silent.  Files are read once and
    kept on the device until they change on disk."""
    if filt in _bpm_registry:
        return _to_dev(_bpm_registry[filt], torch.uint8)
    name = get_par(set_bb.bad_pixel_mask, tel)
    if not isinstance(name, str):
        log.warning('no bad pixel mask configured for telescope %s (set_bb.bad_pixel_mask)', tel)
        return None
    fits_bpm = name.replace('bpm', 'bpm_{}'.format(filt))
    present, fits_bpm = fitsio.already_exists(fits_bpm, get_filename=True)
    if not present:
        log.warning('bad pixel mask %s does not exist', fits_bpm)
        return None
    stat = os.stat(fits_bpm)
    hit = _bpm_cache.get(fits_bpm)
    if hit is not None and hit[0] == stat.st_mtime_ns and hit[1] == stat.st_size and hit[2].device == device:
        return hit[2]
    _, bpm_t = read_fits_image(fits_bpm, dtype=torch.uint8)
    log.info('using bad pixel mask %s', fits_bpm)
    _bpm_cache[fits_bpm] = (stat.st_mtime_ns, stat.st_size, bpm_t)
    return bpm_t


def mask_init(data, header, filt, imgtype, bpm=None):
    """Initial mask from the bad-pixel mask, non-finite pixels, per-channel saturation,
    crosstalk victims, saturated-connected pixels and filled holes
    (blackbox.py:4375-4579, 4584-4596).  Returns (uint8 mask, header_mask dict); ``data``
    has its non-finite pixels zeroed in place.  Uses the module-global ``tel``.  The bad-pixel
    mask is read from ``set_bb.bad_pixel_mask`` as the reference does (see ``_load_bpm``) unless
    ``bpm`` is given."""
    is_np = isinstance(data, np.ndarray)
    t = _inplace_target(data, torch.float32, 'mask_init')
    H, W = t.shape
    bpm_t = _to_dev(bpm, torch.uint8) if bpm is not None else _load_bpm(filt, (H, W), t.device)
    if bpm_t is not None and tuple(bpm_t.shape) != (H, W):
        raise ValueError('bad pixel mask shape {} does not match data {}'.format(tuple(bpm_t.shape), (H, W)))
    header_mask = {}
    if imgtype != 'object':
        mask_t = bpm_t.clone() if bpm_t is not None else torch.zeros((H, W), dtype=torch.uint8, device=t.device)
        return (mask_t.cpu().numpy() if is_np else mask_t), header_mask
    nchans = set_bb.ny * set_bb.nx
    biaslevel = np.array([header['BIASM{}'.format(i + 1)] for i in range(nchans)], dtype=np.float64)
    satlevel_chans = (np.array(get_par(set_bb.satlevel, tel)) * np.array(get_par(set_bb.gain, tel))
                      - biaslevel)
    sat_mean = float(np.mean(satlevel_chans))
    _set(header_mask, 'SATURATE', sat_mean, '[e-] mean saturation threshold')
    _set(header, 'SATURATE', sat_mean, '[e-] mean saturation threshold')
    for i in range(nchans):
        key, descr = 'SATLEV{}'.format(i + 1), '[e-] channel {} saturation threshold'.format(i + 1)
        _set(header, key, round(float(satlevel_chans[i]), 1), descr)
        _set(header_mask, key, round(float(satlevel_chans[i]), 1), descr)
    # seed: bpm | non-finite -> bad | saturated (+ marker) through the fused per-pixel kernel
    geom = _reduced_geometry((H, W), tel)
    satlevel_t = torch.from_numpy(satlevel_chans).to(t.device)
    out_img = torch.empty_like(t)
    work = MaskWork(H, W, t.device)
    ok = False
    if MASK_MORPH_SPARSE:
        mask_t = _apply_seed(t, geom, satlevel_t, bpm_t, out_img, work)
        mask_morph_enqueue(mask_t, tel, work, sparse=True)
        nobj, ok = mask_morph_finish(mask_t, tel, work, sparse=True)
    if not ok:                      # seed list overflow / unconverged holes: dense passes
        mask_t = _apply_seed(t, geom, satlevel_t, bpm_t, out_img, None)
        mask_morph_enqueue(mask_t, tel, work, sparse=False)
        nobj, _ = mask_morph_finish(mask_t, tel, work, sparse=False)
    _set(header_mask, 'NOBJ-SAT', nobj, 'number of saturated objects')
    _set(header, 'NOBJ-SAT', nobj, 'number of saturated objects')
    if is_np:
        data[...] = out_img.cpu().numpy()
        return mask_t.cpu().numpy(), header_mask
    t.copy_(out_img)
    return mask_t, header_mask


def _apply_seed(t, geom, satlevel_t, bpm_t, out_img, work):
    g = geom.as_struct()
    out_mask = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    bits = _bits(tel)
    call('bbx_reduce_apply', _ptr(t), 1, C.byref(g), None, None, None, None, None, _ptr(bpm_t),
         _ptr(satlevel_t), C.byref(bits), _ptr(out_img), _ptr(out_mask),
         _ptr(work.seeds) if work is not None else None,
         _ptr(work.seed_count) if work is not None else None,
         int(work.seed_cap) if work is not None else 0, _stream())
    return out_mask


def mask_header(data_mask, header_mask, tel_=None, counts=None):
    """Per-type pixel counts M-*NUM etc. (blackbox.py:4601-4620).  ``tel_``: telescope name
    (default: the module-global ``tel``, as in the reference).  ``counts``: the eight per-bit pixel
    counts if they are already on the host (FramePipeline takes them from the header block);
    otherwise they are counted here (one pass over the mask, synchronises)."""
    if counts is None:
        t = _to_dev(data_mask, torch.uint8)
        dev_counts = torch.zeros(8, dtype=torch.int64, device=t.device)
        call('bbx_mask_counts', _ptr(t), t.numel(), _ptr(dev_counts), _stream())
        counts = dev_counts.cpu().numpy()
    text = {'bad': 'BP', 'edge': 'EP', 'saturated': 'SP', 'saturated-connected': 'SCP',
            'satellite trail': 'STP', 'cosmic ray': 'CRP'}
    mv = get_par(set_bb.mask_value, tel if tel_ is None else tel_)
    for mask_type, short in text.items():
        value = mv[mask_type]
        _set(header_mask, 'M-{}'.format(short), True, '{} pixels included in mask?'.format(mask_type))
        _set(header_mask, 'M-{}VAL'.format(short), value, 'value added to mask for {} pixels'.format(mask_type))
        _set(header_mask, 'M-{}NUM'.format(short), int(counts[int(value).bit_length() - 1]),
             'number of {} pixels'.format(mask_type))


# -------------------------------------------------------------------------------------------
# cosmic rays
# -------------------------------------------------------------------------------------------
class LacosmicWork:
    def __init__(self, H, W, niter, device, info=None):
        self.buf = torch.empty(query('bbx_lacosmic_work_bytes', H, W), dtype=torch.uint8, device=device)
        self.info = info if info is not None else torch.zeros(4 + max(niter, 1), dtype=torch.int64, device=device)


LAC_LAZY, LAC_DENSE, LAC_LAZY_BG, LAC_FUSED, LAC_STATS = 0, 1, 2, 3, 4   # LAC_FUSED / LAC_STATS: after apply_scan_enqueue / apply_stats_enqueue
LAC_STATUS_OVERFLOW, LAC_STATUS_NEED_BG = 1, 2


def lac_retry_mode(status):
    """Mode to repeat a lazy LACosmic call with, given its non-zero status word."""
    return LAC_DENSE if (status & LAC_STATUS_OVERFLOW) else LAC_LAZY_BG


def lacosmic_enqueue(img_t, inmask_t, crmask_t, sigclip, sigfrac, objlim, readnoise, niter,
                     work, readnoise_dev=None, mode=LAC_LAZY):
    """Enqueue detect_cosmics' iterations (include/bbx.h: bbx_lacosmic).  In lazy mode the
    caller must check ``work.info[2]`` afterwards and repeat densely if it is non-zero."""
    H, W = img_t.shape
    if sigclip < 0 or sigfrac < 0:
        mode = LAC_DENSE
    call('bbx_lacosmic', _ptr(img_t), _ptr(inmask_t), _ptr(crmask_t), H, W,
         float(np.float32(sigclip)), float(np.float32(sigfrac)), float(np.float32(objlim)),
         float(np.float32(readnoise)), _ptr(readnoise_dev), int(niter), int(mode), _ptr(work.buf),
         _ptr(work.info), _stream())


def detect_cosmics(indat, inmask=None, sigclip=4.5, sigfrac=0.3, objlim=5.0, gain=1.0,
                   readnoise=6.5, satlevel=65536.0, pssl=0.0, niter=4, sepmed=True,
                   cleantype='meanmask', fsmode='median', psfmodel='gauss', psffwhm=2.5,
                   psfsize=7, psfk=None, psfbeta=4.765, verbose=False, info=None, mode=None):
    """astroscrappy.detect_cosmics (1.0.8 signature) -> (crmask bool, cleanarr float32).

    ``mode``: None = lazy evaluation with an automatic repeat when its status word asks for it
    (with the background level, or densely: same result either way), LAC_LAZY / LAC_DENSE /
    LAC_LAZY_BG to force one implementation (tests).

    Implemented path: sepmed=False, cleantype='medmask', fsmode='median', pssl=0,
    satlevel=inf -- the one blackbox.py:4323-4332 uses; anything else raises
    NotImplementedError rather than silently computing something different."""
    if sepmed or cleantype != 'medmask' or fsmode != 'median' or pssl != 0.0:
        raise NotImplementedError('detect_cosmics: only sepmed=False, cleantype="medmask", '
                                  'fsmode="median", pssl=0 are implemented')
    if np.isfinite(satlevel):
        raise NotImplementedError('detect_cosmics: finite satlevel is not implemented '
                                  '(the reference passes satlevel=inf, blackbox.py:4272)')
    is_np = isinstance(indat, np.ndarray)
    src = _to_dev(indat, torch.float32)
    inmask_t = None
    if inmask is not None:
        inmask_t = _to_dev(inmask)
        inmask_t = (inmask_t != 0).to(torch.uint8) if inmask_t.dtype != torch.uint8 else inmask_t
        if tuple(inmask_t.shape) != tuple(src.shape):
            raise ValueError('detect_cosmics: inmask shape {} does not match indat {}'.format(
                tuple(inmask_t.shape), tuple(src.shape)))
    if src.dim() != 2:
        raise ValueError('detect_cosmics: indat must be a 2-D image, got shape {}'.format(tuple(src.shape)))
    clean, crmask, work, used_mode, inf, status = _detect_cosmics_dev(
        src, inmask_t, sigclip, sigfrac, objlim, gain, readnoise, niter, mode)
    if info is not None:
        info.update(iterations=int(inf[0]), ncr_per_iter=inf[4:4 + int(inf[0])].copy(),
                    lazy_status=status)
    crb = crmask.to(torch.bool)
    if is_np:
        return crb.cpu().numpy(), clean.cpu().numpy()
    return crb, clean


def _detect_cosmics_dev(src, inmask_t, sigclip, sigfrac, objlim, gain, readnoise, niter, mode):
    """Device part of detect_cosmics: lazy evaluation first, dense repeat if its status word
    asks for it (or if ``mode`` forces one).  Returns (clean, crmask u8, work, mode used,
    info array, lazy status)."""
    H, W = src.shape
    work = LacosmicWork(H, W, niter, src.device)
    crmask = torch.empty((H, W), dtype=torch.uint8, device=src.device)

    def run(m):
        clean = src.clone()
        if gain != 1.0:
            clean *= float(np.float32(gain))
        lacosmic_enqueue(clean, inmask_t, crmask, sigclip, sigfrac, objlim, readnoise, niter, work, mode=m)
        return clean, work.info.cpu().numpy()

    first = LAC_LAZY if mode is None else mode
    if sigclip < 0 or sigfrac < 0:
        first = LAC_DENSE
    clean, inf = run(first)
    status = int(inf[2])
    used = first
    if status != 0:
        if mode is not None:
            raise RuntimeError('detect_cosmics: lazy evaluation incomplete (status {})'.format(status))
        used = lac_retry_mode(status)
        clean, inf = run(used)
        if int(inf[2]) != 0:                 # the lists overflowed on the repeat as well
            used = LAC_DENSE
            clean, inf = run(used)
    if gain != 1.0:
        clean /= float(np.float32(gain))
    return clean, crmask, work, used, inf, status


def cosmics_corr(data, header, data_mask, header_mask):
    """LACosmic detection + cleaning, cosmic-ray bit into the mask, NCOSMICS into both
    headers (blackbox.py:4259-4370).  Uses the module-global ``tel``."""
    if get_par(set_bb.sepmed, tel):
        raise NotImplementedError('cosmics_corr: sepmed=True is not implemented')
    is_np = isinstance(data, np.ndarray)
    d = _to_dev(data, torch.float32)
    m = _to_dev(data_mask, torch.uint8)
    if isinstance(data_mask, torch.Tensor) and m.data_ptr() != data_mask.data_ptr():
        m = m.clone()
    clean, cr8, work, used, _, _ = _detect_cosmics_dev(
        d, m, get_par(set_bb.sigclip, tel), get_par(set_bb.sigfrac, tel), get_par(set_bb.objlim, tel),
        1.0, header['RDNOISE'], get_par(set_bb.niter, tel), None)
    bit = get_par(set_bb.mask_value, tel)['cosmic ray']
    H, W = cr8.shape
    labels = torch.empty(H * W, dtype=torch.int32, device=d.device)
    nobj = torch.zeros(1, dtype=torch.int32, device=d.device)
    call('bbx_lacosmic_finish', _ptr(cr8), _ptr(m), int(bit), H, W, int(used), _ptr(work.buf),
         _ptr(labels), _ptr(nobj), _stream())
    ncosmics_persec = int(nobj.item()) / float(header['EXPTIME'])
    _set(header, 'NCOSMICS', ncosmics_persec, '[/s] number of cosmic rays identified')
    _set(header_mask, 'NCOSMICS', ncosmics_persec, '[/s] number of cosmic rays identified')
    if is_np:
        if isinstance(data_mask, np.ndarray):
            data_mask[...] = m.cpu().numpy()
            return clean.cpu().numpy(), data_mask
        return clean.cpu().numpy(), m.cpu().numpy()
    if isinstance(data_mask, torch.Tensor) and data_mask.data_ptr() != m.data_ptr():
        data_mask.copy_(m)
        m = data_mask
    return clean, m


# -------------------------------------------------------------------------------------------
# non-linearity (blackbox.py:7392-7437; off in the reference's settings)
# -------------------------------------------------------------------------------------------
def spline_tck(spl):
    """(knots, coefficients, degree) of a scipy spline object (UnivariateSpline and its
    subclasses, BSpline) or of a ``(t, c, k)`` tuple."""
    if isinstance(spl, (tuple, list)) and len(spl) == 3:
        t, c, k = spl
    elif hasattr(spl, '_eval_args'):
        t, c, k = spl._eval_args
    elif hasattr(spl, 't') and hasattr(spl, 'c') and hasattr(spl, 'k'):
        t, c, k = spl.t, spl.c, spl.k
    else:
        raise TypeError('nonlin_corr: cannot take knots / coefficients from {!r}'.format(type(spl)))
    return np.asarray(t, dtype=np.float64), np.asarray(c, dtype=np.float64), int(k)


def nonlin_corr(data, nonlin_corr_file, max_counts=50000):
    """In place ``data /= (frac_corr + 1)`` with the per-channel spline of the fractional
    non-linearity evaluated at the counts ``data / gain`` (blackbox.py:7392-7437).
    ``nonlin_corr_file``: path of the pickled list of 16 spline objects the reference reads, or
    that list itself.  Uses the module-global ``tel``.  Returns ``data``."""
    if isinstance(nonlin_corr_file, (str, bytes)):
        import pickle
        with open(nonlin_corr_file, 'rb') as fh:
            splines = pickle.load(fh)
    else:
        splines = nonlin_corr_file
    nchans = set_bb.ny * set_bb.nx
    if len(splines) != nchans:
        raise ValueError('nonlin_corr: {} splines for {} channels'.format(len(splines), nchans))
    tck = [spline_tck(sp) for sp in splines]
    maxn = max(len(t) for t, _, _ in tck)
    knots = np.zeros((nchans, maxn))
    coefs = np.zeros((nchans, maxn))
    for i, (t, c, k) in enumerate(tck):
        knots[i, :len(t)] = t
        coefs[i, :min(len(c), len(t))] = c[:len(t)]
    nk = (C.c_int * nchans)(*[len(t) for t, _, _ in tck])
    deg = (C.c_int * nchans)(*[k for _, _, k in tck])
    is_np = isinstance(data, np.ndarray)
    t_ = _inplace_target(data, torch.float32, 'nonlin_corr')
    H, W = t_.shape
    gain = get_par(set_bb.gain, tel)
    call('bbx_nonlin_corr', _ptr(t_), H, W, H // set_bb.ny, W // set_bb.nx,
         _harr([float(np.float32(g)) for g in gain], C.c_float), knots.ctypes.data_as(C.c_void_p),
         coefs.ctypes.data_as(C.c_void_p), nk, deg, int(maxn), float(max_counts), _stream())
    if is_np:
        data[...] = t_.cpu().numpy()
        return data
    return t_


# -------------------------------------------------------------------------------------------
# FITS data units (fitsio.py reads / writes the files; the byte order is handled here)
# -------------------------------------------------------------------------------------------
def read_fits_image(path, dtype=None):
    """A FITS image file -> (header dict key -> value, CUDA tensor), whatever its packing: an
    uncompressed primary HDU (big-endian bytes decoded by ``bbx_fits_decode``) or an fpacked
    ``.fits.fz`` (Rice-coded tiles decoded by ``bbx_rice_decode``; float images are un-quantised on
    the device).  Stands in for the reference's ``read_hdulist(fits, dtype=...)``
    (blackbox.py:7653-7771) at the places where the hot path reads a file itself: the bad-pixel
    mask of mask_init and the calibration frames of master_prep.  ``dtype``: torch dtype to convert
    to (None: uint16 counts / float32 / uint8 as stored)."""
    with open(path, 'rb') as fh:
        hdr, _ = fitsio.read_header(fh)
    if hdr.get('NAXIS', (0,))[0] == 0:
        ci = fitsio.read_compressed(path, pinned=True)
        hdr = ci.header
        t = rice_decode(ci.heap.to(_device(), non_blocking=True), ci.offsets, ci.lengths, ci.info,
                        zscale=ci.zscale, zzero=ci.zzero, fallback=ci.fallback)
    else:
        hdr, buf, info = fitsio.read_primary(path, pinned=True)
        if info['bitpix'] == 8:
            t = buf.to(_device(), non_blocking=True).view(info['shape'])
        else:
            t = fits_decode(buf.to(_device(), non_blocking=True), info)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return {k: v[0] for k, v in hdr.items()}, t


def fits_decode(be, info, out=None):
    """Big-endian data unit (uint8 CUDA / pinned / numpy buffer as returned by
    ``fitsio.read_primary``) -> native CUDA tensor of shape ``info['shape']``: uint16 counts for
    raw frames (BITPIX 16, BZERO 32768), float32 for BITPIX -32."""
    bitpix, shape = info['bitpix'], tuple(info['shape'])
    t = _to_dev(be if not isinstance(be, np.ndarray) else np.ascontiguousarray(be).view(np.uint8).reshape(-1))
    t = t.view(torch.uint8).reshape(-1)
    n = shape[0] * shape[1]
    if bitpix == 16:
        u16 = info.get('bzero', 0.0) == 32768.0 and info.get('bscale', 1.0) == 1.0
        if not u16 and (info.get('bzero', 0.0) != 0.0 or info.get('bscale', 1.0) != 1.0):
            raise NotImplementedError('fits_decode: BITPIX 16 with BZERO {} / BSCALE {}'.format(
                info.get('bzero'), info.get('bscale')))
        dt = torch.uint16 if u16 else torch.int16
    elif bitpix == -32:
        u16, dt = False, torch.float32
    else:
        raise NotImplementedError('fits_decode: BITPIX {}'.format(bitpix))
    if t.numel() != n * abs(bitpix) // 8:
        raise ValueError('fits_decode: {} bytes for shape {} / BITPIX {}'.format(t.numel(), shape, bitpix))
    if out is None:
        out = torch.empty(shape, dtype=dt, device=t.device)
    call('bbx_fits_decode', _ptr(t), int(bitpix), int(u16), n, _ptr(out), _stream())
    return out


_RICE_DTYPES = {1: torch.uint8, 2: torch.int16, 4: torch.int32}


def rice_decode(heap, offsets, lengths, info, out=None, check=True, zscale=None, zzero=None, fallback=None):
    """Tile-compressed image (``fitsio.read_compressed``: heap bytes + per-tile descriptors) ->
    native CUDA tensor of shape ``info['shape']``: uint16 counts for BITPIX 16 / BZERO 32768 (what
    read_hdulist returns for an fpacked raw frame, blackbox.py:1451), int16 / uint8 / int32 for the
    other integer images, float32 for quantised float images (``zscale`` / ``zzero``: the table's
    per-tile columns).  The heap crosses PCIe compressed; ``bbx_rice_decode`` unpacks it.
    ``check``: synchronise and raise on a corrupt tile (pass False inside a pipeline and test the
    returned status later).  ``fallback``: {row: values} of the tiles that are not Rice-coded
    (``CompressedImage.fallback``; their length is 0 and the decoder leaves the rows alone)."""
    shape = tuple(info['shape'])
    bitpix, bytepix = info.get('bitpix'), info.get('bytepix', 2)
    if (bitpix, bytepix) not in ((8, 1), (16, 2), (32, 4), (-32, 4)):
        raise NotImplementedError('rice_decode: BITPIX {} / BYTEPIX {}'.format(bitpix, bytepix))
    u16 = bitpix == 16 and info.get('bzero', 0.0) == 32768.0 and info.get('bscale', 1.0) == 1.0
    if bitpix != -32 and not u16 and (info.get('bzero', 0.0) != 0.0 or info.get('bscale', 1.0) != 1.0):
        raise NotImplementedError('rice_decode: BZERO {} / BSCALE {}'.format(info.get('bzero'), info.get('bscale')))
    h = _to_dev(heap if not isinstance(heap, np.ndarray) else np.ascontiguousarray(heap)).view(torch.uint8).reshape(-1)
    offs = _to_dev(np.ascontiguousarray(offsets, dtype=np.int64) if not isinstance(offsets, torch.Tensor) else offsets)
    lens = _to_dev(np.ascontiguousarray(lengths, dtype=np.int32) if not isinstance(lengths, torch.Tensor) else lengths)
    if offs.numel() != shape[0] or lens.numel() != shape[0] or offs.dtype != torch.int64 or lens.dtype != torch.int32:
        raise ValueError('rice_decode: {} / {} descriptors for {} tiles'.format(offs.numel(), lens.numel(), shape[0]))
    dt = torch.uint16 if u16 else _RICE_DTYPES[bytepix]
    ints = out if (out is not None and bitpix != -32) else torch.empty(shape, dtype=dt, device=h.device)
    if ints.dtype != dt or tuple(ints.shape) != shape:
        raise ValueError('rice_decode: output must be a {} tensor of shape {}'.format(dt, shape))
    status = torch.empty(1, dtype=torch.int32, device=h.device)
    call('bbx_rice_decode', _ptr(h), h.numel(), _ptr(offs), _ptr(lens), shape[0], shape[1],
         int(info.get('blocksize', 32)), int(bytepix), int(u16), _ptr(ints), _ptr(status), _stream())
    res = ints
    if bitpix == -32:
        if zscale is None or zzero is None:
            raise ValueError('rice_decode: a quantised float image needs its ZSCALE / ZZERO columns')
        method = {'NO_DITHER': 0, 'SUBTRACTIVE_DITHER_1': 1, 'SUBTRACTIVE_DITHER_2': 2}[info.get('quantize') or 'NO_DITHER']
        zs = _to_dev(np.ascontiguousarray(zscale, dtype=np.float64))
        zz = _to_dev(np.ascontiguousarray(zzero, dtype=np.float64))
        rnd = _to_dev(fitsio.dither_random_table()) if method else None
        res = out if out is not None else torch.empty(shape, dtype=torch.float32, device=h.device)
        zblank = info.get('zblank')
        call('bbx_unquantize', _ptr(ints), shape[0], shape[1], _ptr(zs), _ptr(zz), _ptr(rnd), method,
             int(info.get('zdither0', 1) or 1), int(zblank if zblank is not None else 0), int(zblank is not None),
             _ptr(res), _stream())
    empty = (np.asarray(lengths.cpu() if isinstance(lengths, torch.Tensor) else lengths) == 0).nonzero()[0]
    if set(int(r) for r in empty) - set(fallback or {}):
        raise ValueError('rice_decode: {} tile(s) without bytes and without fall-back values'.format(len(empty)))
    for r, vals in (fallback or {}).items():
        res[int(r)].copy_(torch.from_numpy(np.ascontiguousarray(vals)).to(res.dtype))
    if check:
        code = int(status.item())
        if code:
            raise ValueError('rice_decode: corrupt compressed tile(s), status {}'.format(code))
        return res
    return res, status


class RiceEncoder:
    """Scratch and output buffers of ``bbx_rice_encode`` for images of one shape (reusable).
    ``out_bytes``: size of the device output buffer (descriptors + heap); None = the size that
    always fits.  The mask compresses ~50-fold, so FramePipeline asks for a few MB."""

    def __init__(self, shape, bytepix, device, out_bytes=None):
        self.shape, self.bytepix = tuple(shape), int(bytepix)
        H, W = self.shape
        self.work = torch.empty(query('bbx_rice_encode_work_bytes', H, W, self.bytepix), dtype=torch.uint8, device=device)
        full = query('bbx_rice_encode_out_bytes', H, W, self.bytepix)
        self.heap_offset = 16 + (4 * H + 15) // 16 * 16
        self.out_bytes = int(full if out_bytes is None else max(int(out_bytes), self.heap_offset + 16))
        self.out = torch.empty(self.out_bytes, dtype=torch.uint8, device=device)

    def enqueue(self, img_t):
        H, W = self.shape
        if tuple(img_t.shape) != self.shape or img_t.element_size() != self.bytepix or not img_t.is_contiguous():
            raise ValueError('rice_encode: expected a contiguous {}-byte image of shape {}'.format(self.bytepix, self.shape))
        call('bbx_rice_encode', _ptr(img_t), H, W, self.bytepix, _ptr(self.work), self.work.numel(),
             _ptr(self.out), self.out.numel(), _stream())
        return self.out

    def parse(self, host_buf):
        """(total heap bytes, lengths int32 [H], heap view, fits) of an output buffer copied to the
        host (numpy uint8 / pinned tensor).  ``fits`` False: the heap was larger than the buffer."""
        raw = host_buf.numpy() if hasattr(host_buf, 'numpy') else np.asarray(host_buf)
        total = int(raw[0:8].view(np.int64)[0])
        status = int(raw[12:16].view(np.int32)[0])
        H = self.shape[0]
        lens = raw[16:16 + 4 * H].view(np.int32)
        fits = status == 0 and self.heap_offset + total <= raw.size
        return total, lens, raw[self.heap_offset:self.heap_offset + min(total, raw.size - self.heap_offset)], fits


class FpackEncoder:
    """Scratch and output buffers of ``bbx_fpack_f32`` for float32 images of one shape: what the
    reference's ``fpack -q 16 -D -Y`` makes of a reduced image (blackbox.py:826-836), made on the
    device.  ``out_bytes``: size of the output buffer (table columns + heap); None = always fits."""

    def __init__(self, shape, device, q=16.0, out_bytes=None):
        self.shape, self.q = tuple(shape), float(q)
        H, W = self.shape
        self.work = torch.empty(query('bbx_fpack_f32_work_bytes', H, W), dtype=torch.uint8, device=device)
        self.heap_offset = int(query('bbx_fpack_f32_heap_offset', H))
        full = query('bbx_fpack_f32_out_bytes', H, W)
        self.out_bytes = int(full if out_bytes is None else max(int(out_bytes), self.heap_offset + 16))
        self.out = torch.empty(self.out_bytes, dtype=torch.uint8, device=device)
        self.rand = _to_dev(fitsio.dither_random_table())

    def enqueue(self, img_t, zdither0=1):
        H, W = self.shape
        if tuple(img_t.shape) != self.shape or img_t.dtype != torch.float32 or not img_t.is_contiguous():
            raise ValueError('fpack_f32: expected a contiguous float32 image of shape {}'.format(self.shape))
        call('bbx_fpack_f32', _ptr(img_t), H, W, self.q, int(zdither0), _ptr(self.rand), _ptr(self.work),
             self.work.numel(), _ptr(self.out), self.out.numel(), _stream())
        return self.out

    def parse(self, host_buf):
        """-> dict(total, lengths int32 [H], zscale / zzero float64 [H], heap view, fits, skipped)
        of an output buffer (or its leading part) copied to the host.  ``fits`` False: the buffer
        does not hold the whole heap; ``skipped``: rows that were not quantised (length 0, ZSCALE 0)."""
        raw = host_buf.numpy() if hasattr(host_buf, 'numpy') else np.asarray(host_buf)
        total = int(raw[0:8].view(np.int64)[0])
        status = int(raw[12:16].view(np.int32)[0])
        H = self.shape[0]
        o1 = 16 + (4 * H + 15) // 16 * 16
        o2 = o1 + (8 * H + 15) // 16 * 16
        lens = raw[16:16 + 4 * H].view(np.int32)
        fits = (status & 1) == 0 and self.heap_offset + total <= raw.size
        return dict(total=total, lengths=lens, zscale=raw[o1:o1 + 8 * H].view(np.float64),
                    zzero=raw[o2:o2 + 8 * H].view(np.float64), fits=fits, skipped=status >> 8,
                    heap=raw[self.heap_offset:self.heap_offset + min(total, raw.size - self.heap_offset)])


def fpack_f32(data, q=16.0, zdither0=1):
    """float32 image (numpy or CUDA tensor) -> dict(heap, lengths, zscale, zzero, zdither0,
    lossless_rows) -- the keyword arguments ``fitsio.write_compressed(path, shape=..., zbitpix=-32,
    **packed)`` wants: the image as ``fpack -q <q> -D -Y`` stores it (blackbox.py:836).  Synchronises."""
    t = _to_dev(data)
    if t.dtype != torch.float32 or t.dim() != 2:
        raise NotImplementedError('fpack_f32: dtype {} / {} dimensions'.format(t.dtype, t.dim()))
    t = t.contiguous()
    enc = FpackEncoder(tuple(t.shape), t.device, q)
    got = enc.parse(enc.enqueue(t, zdither0).cpu().numpy())
    if not got['fits']:
        raise RuntimeError('fpack_f32: output buffer too small')
    rows = {int(r): t[int(r)].cpu().numpy() for r in np.nonzero(got['lengths'] == 0)[0]}
    return dict(heap=got['heap'].copy(), lengths=got['lengths'].copy(), zscale=got['zscale'].copy(),
                zzero=got['zzero'].copy(), zdither0=int(zdither0), lossless_rows=rows)


def rice_encode(data):
    """uint8 / int16 / uint16 / int32 image (numpy or CUDA tensor) -> (heap uint8 numpy array,
    lengths int32 [rows]): every row Rice-coded as fits_rcomp_byte / _short / fits_rcomp would
    (uint16 counts are stored as int16 with BZERO 32768, as FITS does).  Feed them to
    ``fitsio.write_compressed``.  Synchronises."""
    t = _to_dev(data)
    if t.dtype == torch.uint16:
        t = (t.view(torch.int16) ^ torch.tensor(-32768, dtype=torch.int16, device=t.device))
    if t.dtype not in (torch.uint8, torch.int16, torch.int32) or t.dim() != 2:
        raise NotImplementedError('rice_encode: dtype {} / {} dimensions'.format(t.dtype, t.dim()))
    enc = RiceEncoder(tuple(t.shape), t.element_size(), t.device)
    host = enc.enqueue(t.contiguous()).cpu().numpy()
    total, lens, heap, fits = enc.parse(host)
    if not fits:
        raise RuntimeError('rice_encode: output buffer too small')
    return heap.copy(), lens.copy()


def fits_encode(data, out=None):
    """Native CUDA tensor (float32, uint16, int16 or uint8) -> big-endian data unit as a uint8
    CUDA tensor (uint16 is stored as int16 with BZERO 32768; uint8 needs no swap)."""
    t = _to_dev(data)
    n = t.numel()
    if t.dtype == torch.uint8:
        return t.reshape(-1), 8
    if t.dtype == torch.float32:
        bitpix, u16 = -32, 0
    elif t.dtype in (torch.uint16, torch.int16):
        bitpix, u16 = 16, int(t.dtype == torch.uint16)
    else:
        raise NotImplementedError('fits_encode: dtype {}'.format(t.dtype))
    if out is None:
        out = torch.empty(n * abs(bitpix) // 8, dtype=torch.uint8, device=t.device)
    call('bbx_fits_encode', _ptr(t), bitpix, u16, n, _ptr(out), _stream())
    return out, bitpix


# -------------------------------------------------------------------------------------------
# edge pixels (blackbox.py:1958-1974)
# -------------------------------------------------------------------------------------------
def channel_medians(data, ignore_nan=False):
    """np.median (``ignore_nan``: np.nanmedian, as get_flatstats takes it, blackbox.py:3728-3733)
    of each of the 16 channels of a reduced frame -> float32 [16] CUDA tensor."""
    t = _to_dev(data, torch.float32)
    H, W = t.shape
    ny, nx = set_bb.ny, set_bb.nx
    if H % ny or W % nx:
        raise ValueError('frame {} is not {} x {} channels'.format(tuple(t.shape), ny, nx))
    work = torch.empty(query('bbx_chanmed_work_bytes'), dtype=torch.uint8, device=t.device)
    med = torch.empty(ny * nx, dtype=torch.float32, device=t.device)
    call('bbx_channel_medians', _ptr(t), H, W, H // ny, W // nx, int(bool(ignore_nan)), _ptr(work), _ptr(med),
         _stream())
    return med


def fill_edge_pixels(data, data_mask, medians=None):
    """In place: every 'edge' pixel becomes the median of its channel ("to avoid initial
    source-extractor run leading to a wrong background estimation near the edge",
    blackbox.py:1958-1974).  Uses the module-global ``tel``.  Returns the channel medians."""
    is_np = isinstance(data, np.ndarray)
    t = _inplace_target(data, torch.float32, 'fill_edge_pixels')
    m = _to_dev(data_mask, torch.uint8)
    if tuple(m.shape) != tuple(t.shape):
        raise ValueError('mask shape {} does not match data {}'.format(tuple(m.shape), tuple(t.shape)))
    med = channel_medians(t) if medians is None else _to_dev(medians, torch.float32)
    H, W = t.shape
    edge = int(get_par(set_bb.mask_value, tel)['edge'])
    call('bbx_fill_edge', _ptr(t), _ptr(m), H, W, H // set_bb.ny, W // set_bb.nx, edge, _ptr(med), _stream())
    if is_np:
        data[...] = t.cpu().numpy()
    return med


# -------------------------------------------------------------------------------------------
# crosstalk
# -------------------------------------------------------------------------------------------
def read_crosstalk_file(crosstalk_file, nchans=16):
    """ASCII table with columns victim / source / correction (1-based channel numbers, column
    names on the first line; blackbox.py:7157-7161) -> float64 [source, victim] matrix."""
    coeffs = np.zeros((nchans, nchans))
    with open(crosstalk_file) as fh:
        rows = [ln.split() for ln in fh if ln.strip() and not ln.lstrip().startswith('#')]
    names = ['victim', 'source', 'correction']
    try:
        float(rows[0][0])
    except ValueError:
        names = [n.lower() for n in rows[0]]
        rows = rows[1:]
    iv, isrc, ic = names.index('victim'), names.index('source'), names.index('correction')
    for r in rows:
        coeffs[int(float(r[isrc])) - 1, int(float(r[iv])) - 1] = float(r[ic])
    return coeffs


def xtalk_enqueue(img_t, mask_t, coeffs, tel_, counts=None, variant=0):
    """Enqueue the crosstalk correction.  ``counts`` (int64 [8] device tensor): also receives the
    pixels per mask bit (the kernel sees every mask byte anyway).  ``variant``: 0 = the kernel
    bbx_xtalk picks, 1..4 see include/bbx.h."""
    H, W = img_t.shape
    bits = _bits(tel_)
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    call('bbx_xtalk_counts', _ptr(img_t), _ptr(mask_t), H, W, H // 2, W // 8,
         c.ctypes.data_as(C.c_void_p), C.byref(bits), int(variant), _ptr(counts), _stream())


def xtalk_corr(data, crosstalk_file, data_mask=None):
    """In-place crosstalk correction (blackbox.py:7138-7258).  ``crosstalk_file`` is the path
    of the coefficient table or an already parsed [source, victim] matrix.  ``data_mask`` is
    left unchanged.  Uses the module-global ``tel``."""
    coeffs = crosstalk_file if isinstance(crosstalk_file, np.ndarray) else read_crosstalk_file(crosstalk_file)
    is_np = isinstance(data, np.ndarray)
    t = _inplace_target(data, torch.float32, 'xtalk_corr')
    m = None
    if data_mask is not None:
        m = _to_dev(data_mask)
        if m.dtype != torch.uint8:
            m = m.to(torch.uint8)
    if t.shape[0] % 2 or t.shape[1] % 8:
        raise ValueError('xtalk_corr: frame {} is not 2 x 8 channels'.format(tuple(t.shape)))
    if m is not None and tuple(m.shape) != tuple(t.shape):
        raise ValueError('xtalk_corr: mask shape {} does not match data {}'.format(tuple(m.shape), tuple(t.shape)))
    if np.shape(coeffs) != (16, 16):
        raise ValueError('xtalk_corr: coefficient matrix has shape {}, expected (16, 16)'.format(np.shape(coeffs)))
    xtalk_enqueue(t, m, coeffs, tel)
    if is_np:
        data[...] = t.cpu().numpy()


# -------------------------------------------------------------------------------------------
# master frames
# -------------------------------------------------------------------------------------------
def master_combine(frames, imgtype='bias', medsec=None, bpm=None, tel=None, out=None, clip_sigma=None,
                   clip_maxiters=5, out_ptrs=None, multicast=False):
    """Arithmetic core of master_prep (blackbox.py:4908-4984, 5063-5073): per-pixel median of
    the stack; flats are first divided by their normalisation median (``medsec[i]`` = the
    header's MEDSEC, else the median over set_bb.flat_norm_sec) and get edge / non-positive
    pixels set to 1 afterwards.  ``frames``: sequence of float32 arrays / CUDA tensors of one
    shape (or a 3-D array).  Returns (master, scales).

    ``out_ptrs``: raw device addresses (ints) of up to 8 buffers that all receive the result -- this
    GPU's and its peers' (``distributed.PeerMaster``); ``multicast``: ``out_ptrs[0]`` is an NVSwitch
    multicast address instead.  ``out`` is then only what is returned.

    ``clip_sigma`` (default None = the reference's plain median): sigma-clip every pixel's stack
    first (astropy.stats.sigma_clip, cenfunc='median', ``clip_maxiters`` rounds) and take the
    median of what is left -- the combine BASELINE.json words; not what master_prep does."""
    is_np = isinstance(frames[0], np.ndarray)
    ts = [_to_dev(f, torch.float32) for f in frames]
    n = len(ts)
    if n < 1 or n > 64:
        raise ValueError('master_combine: {} frames (supported: 1..64)'.format(n))
    shape = tuple(ts[0].shape)
    for i, t in enumerate(ts):
        if tuple(t.shape) != shape:
            raise ValueError('master_combine: frame {} has shape {}, frame 0 has {}'.format(i, tuple(t.shape), shape))
    if imgtype == 'flat' and medsec is not None and len(medsec) != n:
        raise ValueError('master_combine: {} normalisation medians for {} frames'.format(len(medsec), n))
    if bpm is not None and imgtype == 'flat' and tuple(np.shape(bpm)) != shape:
        raise ValueError('master_combine: bad-pixel mask shape {} does not match the frames {}'.format(
            tuple(np.shape(bpm)), shape))
    scales = [0.0] * n
    if imgtype == 'flat':
        sec = get_par(set_bb.flat_norm_sec, tel)
        for i, t in enumerate(ts):
            if medsec is not None and medsec[i] is not None:
                med = float(medsec[i])
            else:
                med = float(exact_median(t[sec]))
            scales[i] = med
    bpm_t = _to_dev(bpm, torch.uint8) if (imgtype == 'flat' and bpm is not None) else None
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=ts[0].device)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    scale_h = _harr([float(np.float32(s)) for s in scales], C.c_float)
    flat_fix = 1 if (imgtype == 'flat' and bpm_t is not None) else 0
    edge = int(get_par(set_bb.mask_value, tel)['edge'])
    if out_ptrs is not None:
        if clip_sigma is not None:
            raise ValueError('master_combine: the clipped combine writes to one buffer')
        dsts = (C.c_void_p * len(out_ptrs))(*[int(p) for p in out_ptrs])
        call('bbx_stack_median_multi', ptrs, scale_h, n, out.numel(), flat_fix, _ptr(bpm_t), edge, dsts,
             len(out_ptrs), int(bool(multicast)), _stream())
    elif clip_sigma is None:
        call('bbx_stack_median', ptrs, scale_h, n, out.numel(), flat_fix, _ptr(bpm_t), edge, _ptr(out), _stream())
    else:
        call('bbx_stack_clipped_median', ptrs, scale_h, n, out.numel(), float(clip_sigma), int(clip_maxiters),
             flat_fix, _ptr(bpm_t), edge, _ptr(out), _stream())
    return (out.cpu().numpy() if is_np else out), scales


def exact_median(t):
    """np.median of a float32 CUDA tensor: mean of the two middle values for even counts
    (float32), via two exact rank selections."""
    flat = t.contiguous().view(-1)
    n = flat.numel()
    if n == 0:
        return float('nan')
    srt = torch.sort(flat).values
    if n % 2:
        return srt[n // 2].item()
    a, b = srt[n // 2 - 1], srt[n // 2]
    return ((a + b) / 2).item()
