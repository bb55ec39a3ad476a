"""CCD channel geometry: the layout contract of the hot path.

``define_sections`` returns the same five tuples of 16 ``(yslice, xslice)`` pairs as the
reference (blackbox.py:6334-6402).  ``Geometry`` condenses them into the plain integers the
CUDA kernels take (``bbx_geom`` in include/bbx.h); every kernel is shape-parametric through it.

Channel index = 8*row + col with row 0 the bottom half of the CCD (blackbox.py:6356-6367).
"""
import ctypes
from dataclasses import dataclass

from . import set_bb
from .set_bb import get_par


def define_sections(data_shape, xbin=1, ybin=1, tel=None):
    """(chan_sec, data_sec, os_sec_hori, os_sec_vert, data_sec_red) for a raw or reduced
    frame of ``data_shape``; reference: blackbox.py:6334-6402."""
    ysize, xsize = data_shape
    ny = get_par(set_bb.ny, tel)
    nx = get_par(set_bb.nx, tel)
    dy, dx = ysize // ny, xsize // nx
    ysize_chan = get_par(set_bb.ysize_chan, tel) // ybin
    xsize_chan = get_par(set_bb.xsize_chan, tel) // xbin
    ysize_os = (ysize - ny * ysize_chan) // ny
    xsize_os = (xsize - nx * xsize_chan) // nx

    xs = range(0, xsize, dx)
    chan_sec = tuple((slice(y, y + dy), slice(x, x + dx))
                     for y in range(0, ysize, dy) for x in xs)
    data_sec = tuple((slice(y, y + ysize_chan), slice(x, x + xsize_chan))
                     for y in range(0, ysize, dy + ysize_os) for x in xs)
    # vertical overscan: skip the first few columns (image flux leaks in) and the last one
    ncut_vert = max(5 // xbin, 1)
    os_sec_vert = tuple((slice(y, y + dy), slice(x + xsize_chan + ncut_vert, x + dx - 1))
                        for y in range(0, ysize, dy) for x in xs)
    # horizontal overscan: drop the rows nearest the image
    ncut_hori = max(10 // ybin, 1)
    ysize_os_cut = ysize_os - ncut_hori
    os_sec_hori = tuple((slice(y, y + ysize_os_cut), slice(x, x + dx))
                        for y in range(dy - ysize_os_cut, dy + ysize_os_cut, ysize_os_cut)
                        for x in xs)
    data_sec_red = tuple((slice(y, y + ysize_chan), slice(x, x + xsize_chan))
                         for y in range(0, ysize - ny * ysize_os, ysize_chan)
                         for x in range(0, xsize - nx * xsize_os, xsize_chan))
    return chan_sec, data_sec, os_sec_hori, os_sec_vert, data_sec_red


class BbxGeom(ctypes.Structure):
    """ctypes twin of ``bbx_geom`` (include/bbx.h)."""
    _fields_ = [(n, ctypes.c_int) for n in (
        'H', 'W', 'ny', 'nx', 'dy', 'dx', 'ysize_chan', 'xsize_chan',
        'vos_x0', 'vos_w', 'hos_rows',
        'data_y0_bot', 'data_y0_top', 'hos_y0_bot', 'hos_y0_top')]


@dataclass(frozen=True)
class Geometry:
    """Integer description of a raw frame (with overscans) for the kernels."""
    H: int
    W: int
    ny: int
    nx: int
    dy: int            # channel tile height incl. horizontal overscan
    dx: int            # channel tile width incl. vertical overscan
    ysize_chan: int    # data rows per channel
    xsize_chan: int    # data columns per channel
    vos_x0: int        # first column of the vertical-overscan strip, relative to the tile
    vos_w: int         # width of that strip
    hos_rows: int      # rows of the horizontal-overscan strip
    data_y0: tuple     # first raw row of the data section, per CCD half (bottom, top)
    hos_y0: tuple      # first raw row of the horizontal-overscan strip, per CCD half

    @property
    def nchans(self):
        return self.ny * self.nx

    @property
    def red_shape(self):
        return (self.ny * self.ysize_chan, self.nx * self.xsize_chan)

    @classmethod
    def from_raw_shape(cls, data_shape, xbin=1, ybin=1, tel=None):
        chan_sec, data_sec, os_h, os_v, _ = define_sections(data_shape, xbin, ybin, tel)
        ny = get_par(set_bb.ny, tel)
        nx = get_par(set_bb.nx, tel)
        if ny != 2 or len(chan_sec) != ny * nx or len(data_sec) != ny * nx \
                or len(os_h) != ny * nx or len(os_v) != ny * nx:
            raise ValueError('shape {} does not describe a 2x{} channel raw frame with '
                             'overscans'.format(tuple(data_shape), nx))
        H, W = data_shape
        dy = chan_sec[0][0].stop - chan_sec[0][0].start
        dx = chan_sec[0][1].stop - chan_sec[0][1].start
        ysc = data_sec[0][0].stop - data_sec[0][0].start
        xsc = data_sec[0][1].stop - data_sec[0][1].start
        vos_x0 = os_v[0][1].start
        vos_w = os_v[0][1].stop - os_v[0][1].start
        hos_rows = os_h[0][0].stop - os_h[0][0].start
        if vos_w <= 0 or hos_rows <= 0:
            raise ValueError('raw frame {} has no usable overscan sections'
                             .format(tuple(data_shape)))
        return cls(H, W, ny, nx, dy, dx, ysc, xsc, vos_x0, vos_w, hos_rows,
                   (data_sec[0][0].start, data_sec[nx][0].start),
                   (os_h[0][0].start, os_h[nx][0].start))

    def as_struct(self):
        return BbxGeom(self.H, self.W, self.ny, self.nx, self.dy, self.dx,
                       self.ysize_chan, self.xsize_chan, self.vos_x0, self.vos_w,
                       self.hos_rows, self.data_y0[0], self.data_y0[1],
                       self.hos_y0[0], self.hos_y0[1])
