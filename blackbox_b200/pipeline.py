"""Device-resident reduction chain for one frame after the other (the batched hot path).

``FramePipeline.enqueue(raw)`` puts the whole chain of blackbox_reduce's array steps
(blackbox.py:1479-1902: gain -> overscan -> master bias -> mask_init -> master flat ->
LACosmic -> crosstalk) on the current CUDA stream.  Nothing is allocated per frame: every
scratch buffer is created once per pipeline.  There is exactly one short host round trip per
frame, right after the overscan statistics (< 1 ms of GPU work): the host reads the 16
"spline needed" flags and, for the channels that have saturated-star columns among the first
150 horizontal-overscan columns, evaluates FITPACK's smoothing spline (hostfit.py).
``finish()`` synchronises, checks the remaining device status word (hole filling that needs
more rounds) and returns the header values.

Frames are independent, so a night batch shards one frame per GPU (``shard_frames``); nothing
is exchanged between ranks.
"""
import numpy as np
import torch

from . import reduce as R
from . import set_bb
from ._lib import call
from .geometry import Geometry
from .set_bb import get_par


class FrameResult:
    """Outputs and header values of one reduced frame."""

    def __init__(self, img, mask, header, header_mask, spline_columns=0, redo=False):
        self.img, self.mask = img, mask
        self.header, self.header_mask = header, header_mask
        self.spline_columns = spline_columns   # overscan columns taken from the host spline
        self.redo = redo                       # hole filling needed extra rounds -> chain redone


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
        self._flags_host = torch.zeros(2 * self.geom.nchans, dtype=torch.uint8).pin_memory()
        self._raw = None
        self._out = None
        self._spline_cols = 0

    # ---------------------------------------------------------------------------------------
    def _gain_for(self, raw_t):
        return self.gain if R._raw_type(raw_t) == 0 else None

    def _overscan(self, raw_t):
        """Overscan statistics + fits on the device, then the one host round trip."""
        st = self.st
        R.overscan_enqueue(raw_t, self.geom, self.tel, gain=self._gain_for(raw_t), state=st)
        flags = torch.cat([st.need_spline.any(dim=1).to(torch.uint8),
                           (st.fit_status != 0).to(torch.uint8)])
        self._flags_host.copy_(flags, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        f = self._flags_host.numpy()
        n = self.geom.nchans
        self._spline_cols = 0
        if f[n:].any() or f[:n].any():
            self._spline_cols = R.overscan_resolve_spline(st, strict=False)
        call('bbx_header_means', R._ptr(st.biasm), R._ptr(st.std_vos), R._ptr(self.means), R._stream())

    def _rest(self, raw_t, out_img, out_mask, dense_morph=False, lac_mode=R.LAC_LAZY):
        tel, geom = self.tel, self.geom
        RH, RW = geom.red_shape
        s = R._stream()
        R.apply_enqueue(raw_t, geom, tel, st=self.st, gain=self._gain_for(raw_t), mbias=self.mbias,
                        mflat=self.mflat, bpm=self.bpm, want_mask=True, out_img=out_img, out_mask=out_mask,
                        mwork=None if dense_morph else self.mwork)
        R.mask_morph_enqueue(out_mask, tel, self.mwork, count_objects=self.count_objects,
                             sparse=not dense_morph)
        if dense_morph:
            R.mask_morph_finish(out_mask, tel, self.mwork, sparse=False)
        if self.niter > 0:
            R.lacosmic_enqueue(out_img, out_mask, self.crmask, get_par(set_bb.sigclip, tel),
                               get_par(set_bb.sigfrac, tel), get_par(set_bb.objlim, tel), 0.0,
                               self.niter, self.lwork, readnoise_dev=self.means[1:], mode=lac_mode)
            bit = int(get_par(set_bb.mask_value, tel)['cosmic ray'])
            call('bbx_lacosmic_finish', R._ptr(self.crmask), R._ptr(out_mask), bit, RH, RW, int(lac_mode),
                 R._ptr(self.lwork.buf), R._ptr(self.mwork.labels), R._ptr(self.ncosmic), s)
        if self.coeffs is not None:
            R.xtalk_enqueue(out_img, out_mask, self.coeffs, tel)

    def enqueue(self, raw_t, out_img=None, out_mask=None):
        """Run the overscan stage and enqueue the rest of the chain for one raw frame (uint16
        or float32 CUDA tensor).  Returns the output tensors (valid after ``finish``)."""
        RH, RW = self.geom.red_shape
        if out_img is None:
            out_img = torch.empty((RH, RW), dtype=torch.float32, device=self.device)
        if out_mask is None:
            out_mask = torch.empty((RH, RW), dtype=torch.uint8, device=self.device)
        self._overscan(raw_t)
        self._rest(raw_t, out_img, out_mask)
        self._raw, self._out = raw_t, (out_img, out_mask)
        return out_img, out_mask

    def enqueue_status(self, host_slot):
        """Enqueue a copy of the frame's status (hole filling unconverged | lazy LACosmic
        incomplete) into the pinned int32[1] tensor ``host_slot`` (valid after the stream is
        synchronised): non-zero means the frame has to be finished with ``finish()`` before its
        outputs are used."""
        stat = self.mwork.status[0:1]
        if self.niter > 0:
            stat = stat | self.lwork.info[2:3].to(torch.int32)
        host_slot.copy_(stat, non_blocking=True)

    # ---------------------------------------------------------------------------------------
    def finish(self, fill_header=True):
        """Synchronise, verify the device status of the last enqueued frame and return its
        FrameResult."""
        st = self.st
        out_img, out_mask = self._out
        redo = False
        morph_bad = int(self.mwork.status[0].item()) != 0
        lac_status = int(self.lwork.info[2].item()) if self.niter > 0 else 0
        if morph_bad or lac_status != 0:
            # the sparse mask morphology overflowed / did not converge (the mask LACosmic saw was
            # not final) or the lazy LACosmic needs the background level / its dense twin: redo
            # everything after the overscan stage with the kernels concerned
            redo = True
            self._rest(self._raw, out_img, out_mask, dense_morph=morph_bad,
                       lac_mode=R.lac_retry_mode(lac_status) if lac_status != 0 else R.LAC_LAZY)
            torch.cuda.current_stream().synchronize()
            if self.niter > 0 and int(self.lwork.info[2].item()) != 0:
                self._rest(self._raw, out_img, out_mask, dense_morph=morph_bad, lac_mode=R.LAC_DENSE)
                torch.cuda.current_stream().synchronize()
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
        return FrameResult(out_img, out_mask, header, header_mask, self._spline_cols, redo)

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


class BatchReducer:
    """Reduce a batch of raw frames with ``depth`` FramePipelines ping-ponging on their own CUDA
    streams: while the host waits for the overscan flags of one frame (the single round trip of
    the chain) and enqueues its remaining kernels, the GPU is busy with the other frame, so the
    device never idles on the host.  Results are identical to running the frames one by one."""

    def __init__(self, tel, raw_shape, depth=2, **pipeline_kwargs):
        self.depth = int(depth)
        self.pipes = [FramePipeline(tel, raw_shape, **pipeline_kwargs) for _ in range(self.depth)]
        self.streams = [torch.cuda.Stream() for _ in range(self.depth)]

    def run(self, raws, out_imgs, out_masks, fill_header=False):
        """raws: CUDA tensors; out_imgs / out_masks: at least ``depth`` output tensors, frame k
        is written to index k % len(out_imgs).  Returns the FrameResults in order (a frame's
        output buffers are only valid until they are reused)."""
        n, d = len(raws), self.depth
        nout = len(out_imgs)
        if nout < d or len(out_masks) != nout:
            raise ValueError('need at least depth={} output buffers'.format(d))
        results = [None] * n
        caller = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(caller)
        for k in range(n + d):
            j = k % d
            with torch.cuda.stream(self.streams[j]):
                if k >= d:
                    results[k - d] = self.pipes[j].finish(fill_header=fill_header)
                if k < n:
                    self.pipes[j].enqueue(raws[k], out_imgs[k % nout], out_masks[k % nout])
        for s in self.streams:
            caller.wait_stream(s)
        return results
