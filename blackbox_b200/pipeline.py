"""Device-resident reduction chain for one frame after the other (the batched hot path).

``FramePipeline.enqueue(raw)`` puts the whole chain of blackbox_reduce's array steps
(blackbox.py:1479-1902: gain -> overscan -> master bias -> mask_init -> master flat ->
LACosmic -> crosstalk) on the current CUDA stream without any host synchronisation and without
allocating: every scratch buffer is created once per pipeline.  ``finish()`` synchronises,
checks the device-side status words (rarely needed smoothing spline, failed fits, hole filling
that needs more rounds) and returns the header values.

Frames are independent, so a night batch shards one frame per GPU (``shard_frames``); nothing
is exchanged between ranks.
"""
import ctypes as C

import numpy as np
import torch

from . import reduce as R
from . import set_bb
from ._lib import call
from .geometry import Geometry
from .set_bb import get_par


class FrameResult:
    """Outputs and header values of one reduced frame."""

    def __init__(self, img, mask, header, header_mask, redo=False):
        self.img, self.mask = img, mask
        self.header, self.header_mask = header, header_mask
        self.redo = redo            # True if the frame had to take the strict (host spline) path


class FramePipeline:
    def __init__(self, tel, raw_shape, mbias=None, mflat=None, bpm=None, coeffs=None, niter=None,
                 xbin=1, ybin=1, device=None, exptime=60.0, count_objects=True):
        self.tel = tel
        self.device = device if device is not None else R._device()
        self.geom = Geometry.from_raw_shape(tuple(raw_shape), xbin=xbin, ybin=ybin, tel=tel)
        self.gain = [float(x) for x in get_par(set_bb.gain, tel)]
        self.use_bias = bool(get_par(set_bb.subtract_mbias, tel)) and mbias is not None
        self.mbias = R._to_dev(mbias, torch.float32) if self.use_bias else None
        self.mflat = R._to_dev(mflat, torch.float32)
        self.bpm = R._to_dev(bpm, torch.uint8)
        self.coeffs = None if coeffs is None else np.ascontiguousarray(coeffs, dtype=np.float64)
        self.niter = int(get_par(set_bb.niter, tel) if niter is None else niter)
        self.exptime = float(exptime)
        self.count_objects = count_objects
        RH, RW = self.geom.red_shape
        dev = self.device
        self.st = R.OverscanState(self.geom, dev)
        self.mwork = R.MaskWork(RH, RW, dev)
        self.lwork = R.LacosmicWork(RH, RW, self.niter, dev)
        self.crmask = torch.empty((RH, RW), dtype=torch.uint8, device=dev)
        self.means = torch.zeros(2, dtype=torch.float64, device=dev)      # BIASMEAN, RDNOISE
        self.ncosmic = torch.zeros(1, dtype=torch.int32, device=dev)
        self._raw = None
        self._out = None

    # ---------------------------------------------------------------------------------------
    def enqueue(self, raw_t, out_img=None, out_mask=None):
        """Enqueue the full chain for one raw frame (uint16 or float32 CUDA tensor)."""
        tel, geom, st = self.tel, self.geom, self.st
        RH, RW = geom.red_shape
        s = R._stream()
        if out_img is None:
            out_img = torch.empty((RH, RW), dtype=torch.float32, device=self.device)
        if out_mask is None:
            out_mask = torch.empty((RH, RW), dtype=torch.uint8, device=self.device)
        gain = self.gain if R._raw_type(raw_t) == 0 else None
        R.overscan_enqueue(raw_t, geom, tel, gain=gain, state=st)
        call('bbx_header_means', R._ptr(st.biasm), R._ptr(st.std_vos), R._ptr(self.means), s)
        R.apply_enqueue(raw_t, geom, tel, st=st, gain=gain, mbias=self.mbias, mflat=self.mflat,
                        bpm=self.bpm, want_mask=True, out_img=out_img, out_mask=out_mask)
        R.mask_morph_enqueue(out_mask, tel, self.mwork, count_objects=self.count_objects)
        if self.niter > 0:
            R.lacosmic_enqueue(out_img, out_mask, self.crmask, get_par(set_bb.sigclip, tel),
                               get_par(set_bb.sigfrac, tel), get_par(set_bb.objlim, tel), 0.0,
                               self.niter, self.lwork, readnoise_dev=self.means[1:])
            bit = int(get_par(set_bb.mask_value, tel)['cosmic ray'])
            call('bbx_mask_or', R._ptr(out_mask), R._ptr(self.crmask), out_mask.numel(), bit, s)
            if self.count_objects:
                call('bbx_count_objects', R._ptr(self.crmask), 1, RH, RW, R._ptr(self.mwork.labels),
                     R._ptr(self.ncosmic), s)
        if self.coeffs is not None:
            R.xtalk_enqueue(out_img, out_mask, self.coeffs, tel)
        self._raw, self._out = raw_t, (out_img, out_mask)
        return out_img, out_mask

    # ---------------------------------------------------------------------------------------
    def finish(self, fill_header=True):
        """Synchronise, verify the device status of the last enqueued frame and return its
        FrameResult.  Frames that needed the smoothing spline (or more hole-filling rounds)
        are redone through the strict path so the result is always the reference's."""
        st = self.st
        out_img, out_mask = self._out
        redo = False
        need = bool(st.need_spline.any().item())
        if need or bool(st.fit_status.any().item()) or int(self.mwork.unconverged.item()) != 0:
            redo = True
            self._redo_strict()
        header, header_mask = {}, {}
        if fill_header:
            R.fill_os_header(header, st)
            nobj = int(self.mwork.nobj.item())
            header['NOBJ-SAT'] = header_mask['NOBJ-SAT'] = nobj
            sat = st.satlevel.cpu().numpy()
            header['SATURATE'] = header_mask['SATURATE'] = float(np.mean(sat))
            for i in range(self.geom.nchans):
                header['SATLEV{}'.format(i + 1)] = header_mask['SATLEV{}'.format(i + 1)] = round(float(sat[i]), 1)
            if self.niter > 0:
                nc = int(self.ncosmic.item()) / self.exptime
                header['NCOSMICS'] = header_mask['NCOSMICS'] = nc
                info = self.lwork.info.cpu().numpy()
                header['LAC-NIT'] = int(info[0])
        return FrameResult(out_img, out_mask, header, header_mask, redo)

    def _redo_strict(self):
        """Slow path: host spline, then the remaining chain again."""
        tel, geom, st = self.tel, self.geom, self.st
        raw_t = self._raw
        out_img, out_mask = self._out
        gain = self.gain if R._raw_type(raw_t) == 0 else None
        R.overscan_enqueue(raw_t, geom, tel, gain=gain, state=st)
        R.overscan_resolve_spline(st, strict=False)
        s = R._stream()
        call('bbx_header_means', R._ptr(st.biasm), R._ptr(st.std_vos), R._ptr(self.means), s)
        R.apply_enqueue(raw_t, geom, tel, st=st, gain=gain, mbias=self.mbias, mflat=self.mflat,
                        bpm=self.bpm, want_mask=True, out_img=out_img, out_mask=out_mask)
        R.mask_morph_enqueue(out_mask, tel, self.mwork, count_objects=self.count_objects)
        R.mask_morph_finish(out_mask, tel, self.mwork)
        RH, RW = geom.red_shape
        if self.niter > 0:
            R.lacosmic_enqueue(out_img, out_mask, self.crmask, get_par(set_bb.sigclip, tel),
                               get_par(set_bb.sigfrac, tel), get_par(set_bb.objlim, tel), 0.0,
                               self.niter, self.lwork, readnoise_dev=self.means[1:])
            bit = int(get_par(set_bb.mask_value, tel)['cosmic ray'])
            call('bbx_mask_or', R._ptr(out_mask), R._ptr(self.crmask), out_mask.numel(), bit, s)
            if self.count_objects:
                call('bbx_count_objects', R._ptr(self.crmask), 1, RH, RW, R._ptr(self.mwork.labels),
                     R._ptr(self.ncosmic), s)
        if self.coeffs is not None:
            R.xtalk_enqueue(out_img, out_mask, self.coeffs, tel)
        torch.cuda.current_stream().synchronize()

    # ---------------------------------------------------------------------------------------
    def reduce(self, raw):
        """Convenience: numpy or tensor in, FrameResult out (synchronous)."""
        raw_t = R._to_dev(raw)
        self.enqueue(raw_t)
        return self.finish()


def shard_frames(nframes, rank, world_size):
    """Indices of the frames rank ``rank`` reduces: frame k -> GPU k mod world_size
    (frames are independent: the reference runs one process per frame, blackbox.py:378)."""
    return list(range(rank, nframes, world_size))
