"""Geometry contract: define_sections / Geometry against slices derived by hand from the
reference's arithmetic (blackbox.py:6345-6399)."""
import numpy as np
import pytest

from blackbox_b200 import set_bb
from blackbox_b200.geometry import BbxGeom, Geometry, define_sections


def sl(a, b):
    return slice(a, b)


def test_full_raw_frame_sections():
    chan, data, hos, vos, red = define_sections((10600, 12000))
    assert all(len(t) == 16 for t in (chan, data, hos, vos, red))
    assert chan[0] == (sl(0, 5300), sl(0, 1500)) and chan[15] == (sl(5300, 10600), sl(10500, 12000))
    assert data[0] == (sl(0, 5280), sl(0, 1320)) and data[8] == (sl(5320, 10600), sl(0, 1320))
    assert data[15] == (sl(5320, 10600), sl(10500, 11820))
    assert vos[0] == (sl(0, 5300), sl(1325, 1499)) and vos[9] == (sl(5300, 10600), sl(2825, 2999))
    assert hos[0] == (sl(5290, 5300), sl(0, 1500)) and hos[8] == (sl(5300, 5310), sl(0, 1500))
    assert red[0] == (sl(0, 5280), sl(0, 1320)) and red[9] == (sl(5280, 10560), sl(1320, 2640))


def test_binned_frame_sections():
    chan, data, hos, vos, red = define_sections((5300, 6000), xbin=2, ybin=2)
    assert data[0] == (sl(0, 2640), sl(0, 660)) and data[8] == (sl(2660, 5300), sl(0, 660))
    assert vos[0] == (sl(0, 2650), sl(662, 749))
    assert hos[0] == (sl(2645, 2650), sl(0, 750)) and hos[8] == (sl(2650, 2655), sl(0, 750))
    assert red[15] == (sl(2640, 5280), sl(4620, 5280))


def test_reduced_frame_sections_coincide():
    chan, data, hos, vos, red = define_sections((10560, 10560))
    assert chan == data == red
    assert len(hos) == 16 and len(vos) == 16


def test_sections_tile_the_frame():
    for shape, b in [((10600, 12000), 1), ((5300, 6000), 2)]:
        chan, data, hos, vos, red = define_sections(shape, xbin=b, ybin=b)
        cover = np.zeros(shape, np.int8)
        for s in chan:
            cover[s] += 1
        assert (cover == 1).all()
        rshape = (2 * 5280 // b, 8 * 1320 // b)
        cover = np.zeros(rshape, np.int8)
        for s in red:
            cover[s] += 1
        assert (cover == 1).all()
        for d, r in zip(data, red):
            assert (d[0].stop - d[0].start, d[1].stop - d[1].start) == (r[0].stop - r[0].start, r[1].stop - r[1].start)


def test_geometry_struct_matches_sections():
    g = Geometry.from_raw_shape((10600, 12000))
    assert (g.dy, g.dx, g.ysize_chan, g.xsize_chan) == (5300, 1500, 5280, 1320)
    assert (g.vos_x0, g.vos_w, g.hos_rows) == (1325, 174, 10)
    assert g.data_y0 == (0, 5320) and g.hos_y0 == (5290, 5300)
    assert g.red_shape == (10560, 10560) and g.nchans == 16
    s = g.as_struct()
    assert isinstance(s, BbxGeom) and (s.H, s.W, s.hos_y0_top, s.data_y0_top) == (10600, 12000, 5300, 5320)
    gb = Geometry.from_raw_shape((5300, 6000), xbin=2, ybin=2)
    assert (gb.vos_x0, gb.vos_w, gb.hos_rows, gb.data_y0, gb.hos_y0) == (662, 87, 5, (0, 2660), (2645, 2650))


def test_geometry_rejects_frames_without_overscan():
    with pytest.raises(ValueError):
        Geometry.from_raw_shape((10560, 10560))


def test_get_par():
    assert set_bb.get_par(set_bb.sigclip, 'ML1') == 15 and set_bb.get_par(set_bb.sigclip, 'BG3') == 20
    assert set_bb.get_par(set_bb.subtract_mbias, 'BG2') is True and set_bb.get_par(set_bb.subtract_mbias, 'ML1') is False
    assert set_bb.get_par(set_bb.niter, 'BG4') == 3
    assert set_bb.get_par(set_bb.mask_value, 'ML1')['edge'] == 32
    assert len(set_bb.get_par(set_bb.gain, 'BG4')) == 16
