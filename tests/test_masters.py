"""master_prep (blackbox.py:4625-5247): the reference's own function, executed on a synthetic
night held in memory (tests/golden/make_reference_golden.py, 'masters'), against
  * the oracle's stack median + master-flat header statistics (CPU),
  * the host logic of blackbox_b200.masters -- file selection, naming, fall-backs (CPU, the
    combine stubbed out),
  * blackbox_b200.masters.master_prep end to end over FITS files (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'reference_golden.json')))['masters']


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _night(g, with_data=True):
    from blackbox_b200 import set_bb, synth
    shape = (2 * g['ysize_chan'], 8 * set_bb.xsize_chan)
    return shape, synth.make_cal_night(g['tel'], g['imgtype'], g['seed'], shape, g['date_eve'], g['filt'],
                                       with_data=with_data)


def _used(g, night):
    """The frames the reference combined, in its order (header BIAS1.. / FLAT1..)."""
    by_stem = {os.path.basename(n).split('.fits')[0]: (f, h) for n, f, h in night}
    up = g['imgtype'].upper()
    n = g['header']['N' + up]
    return [by_stem[g['header']['{}{}'.format(up, i + 1)]] for i in range(n)]


@pytest.mark.parametrize('idx', range(len(GOLD)))
def test_oracle_master_equals_the_reference(idx, small_bb):
    from blackbox_b200 import synth
    from oracle import reduce as R
    g = GOLD[idx]
    small_bb(g['ysize_chan'])
    shape, night = _night(g)
    used = _used(g, night)
    frames = [f for f, _ in used]
    if g['imgtype'] == 'bias':
        out, _ = R.master_median(frames, imgtype='bias', tel=g['tel'])
        assert digest(out) == g['master_sha256']
        return
    medsec = [h.get('MEDSEC') for _, h in used]
    bpm = synth.make_masters(g['tel'], g['seed'] + 1, shape)[2]
    out, stats = R.master_flat_stats(frames, medsec=medsec, bpm=bpm, tel=g['tel'])
    assert digest(out) == g['master_sha256']
    for key, val in stats.items():
        assert float(val) == g['header'][key], key


def _write_night(root, g, night, data=True):
    from blackbox_b200 import fitsio
    for name, frame, hdr in night:
        path = os.path.join(root, 'red', name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        fitsio.write_primary(path, frame if data else np.zeros((2, 2), np.float32), hdr)


def _fits_master(root, g):
    d = g['date_eve']
    tail = '_' + g['filt'] if g['imgtype'] == 'flat' else ''
    return '{}/masters/{}/{}/{}/{}/{}_{}_{}{}.fits'.format(root, d[0:4], d[4:6], d[6:8], g['imgtype'], g['tel'],
                                                          g['imgtype'], d, tail)


@pytest.fixture
def site(tmp_path, monkeypatch):
    from blackbox_b200 import set_bb
    root = str(tmp_path)
    for name, sub in (('red_dir', 'red'), ('master_dir', 'masters')):
        monkeypatch.setattr(set_bb, name, {t: '{}/{}'.format(root, sub) for t in ('ML1', 'BG2', 'BG3', 'BG4')})
    monkeypatch.setattr(set_bb, 'bad_pixel_mask', {t: '{}/cal/{}_bpm.fits'.format(root, t)
                                                   for t in ('ML1', 'BG2', 'BG3', 'BG4')})
    return root


@pytest.mark.parametrize('idx', range(len(GOLD)))
def test_file_selection_equals_the_reference(idx, site, monkeypatch):
    """Same frames, same order as the reference chose (red flag, evening flats, ncal_max nearest
    to midnight); headers only -- the combine is stubbed, nothing touches the GPU."""
    from blackbox_b200 import masters
    g = GOLD[idx]
    shape, night = _night(g, with_data=False)
    _write_night(site, g, night, data=False)
    seen = {}

    def fake_combine(file_list, data_shape, imgtype, filt, nwindow, tel):
        seen.update(files=file_list, shape=data_shape, imgtype=imgtype, filt=filt, nwindow=nwindow)
        return None, {}

    monkeypatch.setattr(masters, 'combine_files', fake_combine)
    monkeypatch.setattr(masters, 'write_master', lambda name, master, header: name)
    fits_master = _fits_master(site, g)
    assert masters.master_prep(fits_master, shape, True, pick_alt=False, tel=g['tel']) == fits_master
    up = g['imgtype'].upper()
    want = [g['header']['{}{}'.format(up, i + 1)] for i in range(g['header']['N' + up])]
    assert [os.path.basename(f).split('.fits')[0] for f in seen['files']] == want
    assert seen['nwindow'] == g['header'][up + '-WIN'] and seen['shape'] == shape
    assert seen['filt'] == (g['filt'] if g['imgtype'] == 'flat' else None)


def test_master_prep_fallbacks(site, monkeypatch):
    """An existing good master is returned; a red-flagged one is remade; too few frames -> None or
    the nearest master (yesterday's first, else the nearest unflagged one of the month)
    (blackbox.py:4663-4676, 4802-4847, 5294-5395)."""
    from blackbox_b200 import fitsio, masters
    g = GOLD[0]
    shape, night = _night(g, with_data=False)
    tiny = np.zeros((2, 2), np.float32)
    monkeypatch.setattr(masters, 'combine_files', lambda *a: (None, {}))
    made = []
    monkeypatch.setattr(masters, 'write_master', lambda name, m, h: made.append(name) or name)
    fits_master = _fits_master(site, g)
    os.makedirs(os.path.dirname(fits_master), exist_ok=True)

    # too few frames, no alternative
    _write_night(site, g, night[:3], data=False)
    assert masters.master_prep(fits_master, shape, True, pick_alt=False, tel='ML1') is None
    assert masters.master_prep(fits_master, shape, True, pick_alt=True, tel='ML1') is None
    # a master two weeks back and a red-flagged one from three days back: the older good one wins
    old = fits_master.replace('20240105', '20231222').replace('2024/01/05', '2023/12/22')
    bad = fits_master.replace('20240105', '20240102').replace('2024/01/05', '2024/01/02')
    for name, hdr in ((old, {}), (bad, {'QC-FLAG': 'red'})):
        os.makedirs(os.path.dirname(name), exist_ok=True)
        fitsio.write_primary(name, tiny, hdr)
    assert masters.master_prep(fits_master, shape, True, pick_alt=True, tel='ML1') == old
    assert masters.master_prep(fits_master, shape, False, pick_alt=False, tel='ML1') == old
    # yesterday's master is preferred
    yest = fits_master.replace('20240105', '20240104').replace('2024/01/05', '2024/01/04')
    os.makedirs(os.path.dirname(yest), exist_ok=True)
    fitsio.write_primary(yest, tiny, {})
    assert masters.master_prep(fits_master, shape, True, pick_alt=True, tel='ML1') == yest
    assert not made
    # enough frames: made; present and good: returned untouched; present and red: see below
    _write_night(site, g, night, data=False)
    assert masters.master_prep(fits_master, shape, True, pick_alt=False, tel='ML1') == fits_master
    assert made == [fits_master]
    fitsio.write_primary(fits_master, tiny, {})
    assert masters.master_prep(fits_master, shape, True, pick_alt=False, tel='ML1') == fits_master
    assert made == [fits_master]
    fitsio.write_primary(fits_master, tiny, {'QC-FLAG': 'red'})
    # blackbox.py:4802: a red-flagged master is not remade; the nearest good one, if asked for
    assert masters.master_prep(fits_master, shape, True, pick_alt=False, tel='ML1') is None
    assert masters.master_prep(fits_master, shape, True, pick_alt=True, tel='ML1') == yest
    assert made == [fits_master]


def test_all_frames_older_than_12_hours(site, monkeypatch):
    """blackbox.py:4876-4884."""
    from blackbox_b200 import masters
    g = dict(GOLD[0], date_eve='20240109')
    shape, night = _night(GOLD[0], with_data=False)        # frames of 2024-01-04 .. 06
    _write_night(site, g, night, data=False)
    monkeypatch.setattr(masters, 'combine_files', lambda *a: pytest.fail('must not combine'))
    assert masters.master_prep(_fits_master(site, g), shape, True, pick_alt=False, tel='ML1') is None


def test_dates():
    from blackbox_b200 import masters
    assert masters.date2mjd('20240105', '12:00') == 60314.5
    assert masters.date2mjd('2000-01-01') == 51544.0
    assert masters.date2mjd('20240105', '235959') == pytest.approx(60314.0 + 86399 / 86400, abs=1e-9)
    assert masters.mjd2date(60314.5) == '2024/01/05'
    assert [masters.delta_one_month('20240105', d) for d in (-1, 0, 1)] == ['2023/12/', '2024/01/', '2024/02/']
    assert masters.delta_one_month('2024-12-31', 1) == '2025/01/'


@pytest.mark.gpu
@pytest.mark.parametrize('idx', range(len(GOLD)))
def test_master_prep_on_the_gpu_equals_the_reference(idx, site, small_bb):
    from blackbox_b200 import fitsio, masters, set_bb, synth
    g = GOLD[idx]
    small_bb(g['ysize_chan'])
    shape, night = _night(g)
    _write_night(site, g, night)
    if g['imgtype'] == 'flat':
        bpm = synth.make_masters(g['tel'], g['seed'] + 1, shape)[2]
        name = set_bb.get_par(set_bb.bad_pixel_mask, g['tel']).replace('bpm', 'bpm_' + g['filt'])
        os.makedirs(os.path.dirname(name), exist_ok=True)
        fitsio.write_primary(name, bpm)
    fits_master = _fits_master(site, g)
    assert masters.master_prep(fits_master, shape, True, pick_alt=False, tel=g['tel']) == fits_master
    hdr, data, info = fitsio.read_primary(fits_master)
    assert digest(fitsio.to_native(data, info)) == g['master_sha256']
    got = {k: v[0] for k, v in hdr.items()}
    for key, want in g['header'].items():
        if key == 'MFSTDSEC':
            assert got[key] == pytest.approx(want, rel=1e-5)       # float32 np.std: summation order
        elif key == 'OFF-MEAN' or key.startswith('GAINCF') or key in ('RA', 'DEC', 'MJD-OBS', 'MFMEDSEC'):
            assert got[key] == pytest.approx(want, rel=1e-15, abs=0), key   # exact up to the FITS card's repr
        else:
            assert got[key] == want, key
    # a second call finds the master
    assert masters.master_prep(fits_master, shape, True, pick_alt=False, tel=g['tel']) == fits_master


@pytest.mark.gpu
@pytest.mark.parametrize('idx', range(len(GOLD)))
def test_master_prep_over_an_fpacked_night(idx, site, small_bb):
    """What the reference's folders really hold (blackbox.py:826-840: every reduced frame and the
    bad-pixel mask are fpacked): .fits.fz reduced frames -- float images quantised with subtractive
    dithering, Rice-coded -- and an fpacked 8-bit mask.  The master must be the median of the
    frames AS A READER SEES THEM (the oracle's un-quantised values)."""
    from blackbox_b200 import fitsio, masters, set_bb, synth
    from oracle import reduce as R, rice
    g = GOLD[idx]
    small_bb(g['ysize_chan'])
    shape, night = _night(g)
    seen = {}
    for name, frame, hdr in night:
        path = os.path.join(site, 'red', name) + '.fz'
        os.makedirs(os.path.dirname(path), exist_ok=True)
        _, back = rice.write_fz_f32(path, frame, hdr, q=16.0, dither=1, zdither0=1 + len(seen))
        seen[os.path.basename(name).split('.fits')[0]] = (back, hdr)
    bpm = None
    if g['imgtype'] == 'flat':
        bpm = synth.make_masters(g['tel'], g['seed'] + 1, shape)[2]
        name = set_bb.get_par(set_bb.bad_pixel_mask, g['tel']).replace('bpm', 'bpm_' + g['filt'])
        os.makedirs(os.path.dirname(name), exist_ok=True)
        rice.write_fz_u8(name + '.fz', bpm)
    fits_master = _fits_master(site, g)
    assert masters.master_prep(fits_master, shape, True, pick_alt=False, tel=g['tel']) == fits_master
    hdr, data, info = fitsio.read_primary(fits_master)
    got = {k: v[0] for k, v in hdr.items()}
    up = g['imgtype'].upper()
    used = [seen[got['{}{}'.format(up, i + 1)]] for i in range(got['N' + up])]
    assert [got['{}{}'.format(up, i + 1)] for i in range(got['N' + up])] == \
        [g['header']['{}{}'.format(up, i + 1)] for i in range(g['header']['N' + up])]
    frames = [f for f, _ in used]
    if g['imgtype'] == 'bias':
        want, _ = R.master_median(frames, imgtype='bias', tel=g['tel'])
    else:
        want, _ = R.master_flat_stats(frames, medsec=[h.get('MEDSEC') for _, h in used], bpm=bpm, tel=g['tel'])
    assert np.array_equal(fitsio.to_native(data, info), want)


@pytest.mark.gpu
def test_master_prep_degrades_on_an_unreadable_frame(site, small_bb, caplog):
    """A frame this reader cannot unpack (here: GZIP_1 tiles) must not raise out of master_prep: a
    logged error and the nearest existing master, as the reference ends up for unusable input."""
    import logging
    from blackbox_b200 import fitsio, masters
    from oracle import rice
    g = GOLD[0]
    small_bb(g['ysize_chan'])
    shape, night = _night(g)
    for k, (name, frame, hdr) in enumerate(night):
        path = os.path.join(site, 'red', name) + '.fz'
        os.makedirs(os.path.dirname(path), exist_ok=True)
        rice.write_fz_f32(path, frame, hdr)
        if k == 2:
            raw = open(path, 'rb').read().replace(b"'RICE_1  '", b"'GZIP_1  '")
            open(path, 'wb').write(raw)
    fits_master = _fits_master(site, g)
    yest = fits_master.replace('20240105', '20240104').replace('2024/01/05', '2024/01/04')
    os.makedirs(os.path.dirname(yest), exist_ok=True)
    fitsio.write_primary(yest, np.zeros((2, 2), np.float32), {})
    with caplog.at_level(logging.ERROR, logger='blackbox_b200.masters'):
        assert masters.master_prep(fits_master, shape, True, pick_alt=True, tel=g['tel']) == yest
        assert masters.master_prep(fits_master, shape, True, pick_alt=False, tel=g['tel']) is None
    assert any('unreadable input' in r.getMessage() for r in caplog.records)
    assert not os.path.exists(fits_master)
