#!/usr/bin/env python
"""Generate tests/golden/reference_golden.json by EXECUTING THE REFERENCE'S OWN CODE.

/root/reference/blackbox.py cannot be imported as it stands (astropy, astroscrappy, zogy, set_zogy,
fitsio, ephem, watchdog, acstools, ASTA, match2SSO, matplotlib are absent here), but its reduction
functions only need numpy / scipy plus a handful of names from those modules.  This script puts
stub modules in ``sys.modules`` -- empty shells for everything the hot path never touches, and

    astropy.stats.sigma_clipped_stats / sigma_clip   -> oracle.stats      (restated, UNPINNED)
    astroscrappy.detect_cosmics                      -> oracle.lacosmic   (restated, UNPINNED)
    zogy: np, ndimage, interpolate, get_par, fits.Header (a dict), Table.read (3-column ASCII
          reader), read_hdulist / already_exists (in-memory bad-pixel mask)
    set_zogy.mask_value                              -> ZOGY's published defaults

-- imports the reference, and runs define_sections, gain_corr, os_corr, mask_init, cosmics_corr,
xtalk_corr and nonlin_corr UNMODIFIED on seeded synthetic frames.  What is stored are digests,
spot values and header values of the outputs, so tests/test_reference_golden.py can hold the
oracle (and the GPU path) against the reference's own control flow, slicing, dtype promotion,
np.polyfit / UnivariateSpline / ndimage / np.matmul calls.  Only the two stubbed third-party
algorithms stay unpinned.  Needs /root/reference; run here (the fixtures travel, it does not):

    python tests/golden/make_reference_golden.py
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
from scipy import interpolate, ndimage

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference'

USE_REAL_LIBRARIES = '--real' in sys.argv      # default: the oracle's restatements stand in (see the module text)

MASK_VALUE = {'bad': 1, 'cosmic ray': 2, 'saturated': 4, 'saturated-connected': 8,
              'satellite trail': 16, 'edge': 32, 'crosstalk': 64}
_files = {}          # "file name" -> array, served by the read_hdulist stub


class _Sink:
    def __init__(self, name='x'):
        self._n = name

    def __getattr__(self, k):
        return _Sink(self._n + '.' + k)

    def __call__(self, *a, **k):
        return _Sink(self._n + '()')

    def __iter__(self):
        return iter(())


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda k, _n=name: _Sink(_n + '.' + k)
    sys.modules[name] = m
    return m


class Header(dict):
    """astropy.io.fits.Header stand-in: header[key] = (value, comment) stores the value."""

    @property
    def comments(self):
        import collections
        return collections.defaultdict(str)

    def __setitem__(self, key, value):
        if isinstance(value, tuple) and len(value) == 2:
            value = value[0]
        dict.__setitem__(self, key, value)


class _Column:
    def __init__(self, value):
        self.value = value


class _Rows(dict):
    def __len__(self):
        return len(next(iter(self.values())).value)


class _Table:
    """astropy.table.Table.read(..., format='ascii') stand-in: column names on the first line,
    integer columns where every entry is an integer (as astropy's guesser types them)."""

    @staticmethod
    def read(path, format=None, names=None):
        rows = [ln.split() for ln in open(path) if ln.strip() and not ln.lstrip().startswith('#')]
        out = _Rows()
        for i, name in enumerate(rows[0]):
            col = [r[i] for r in rows[1:]]
            try:
                out[name] = _Column(np.array([int(c) for c in col]))
            except ValueError:
                out[name] = _Column(np.array([float(c) for c in col]))
        return out


class Time:
    """astropy.time.Time stand-in for what master_prep needs: Time(iso string or list).mjd,
    Time(mjd, format='mjd').isot, Time.now().isot (UTC; an MJD in UTC involves no leap seconds)."""
    _EPOCH = __import__('datetime').datetime(1858, 11, 17)

    def __init__(self, val, format=None):
        import datetime
        if format == 'mjd':
            self.mjd = float(val)
        elif isinstance(val, (list, tuple)):
            self.mjd = np.array([Time(v).mjd for v in val])
        else:
            d = datetime.datetime.fromisoformat(str(val).strip().replace(' ', 'T')) - self._EPOCH
            self.mjd = d.days + (d.seconds + d.microseconds * 1e-6) / 86400.0

    @property
    def isot(self):
        import datetime
        return (self._EPOCH + datetime.timedelta(days=self.mjd)).isoformat(timespec='milliseconds')

    @classmethod
    def now(cls):
        return cls(60000.0, format='mjd')


def _read_hdulist(name, get_data=True, get_header=False, dtype=None, **k):
    """zogy.read_hdulist stand-in over the in-memory files: entries are arrays or (array, Header)."""
    entry = _files[name]
    data, header = entry if isinstance(entry, tuple) else (entry, Header())
    if get_data and get_header:
        return np.array(data, dtype=dtype, copy=True), header
    if get_header:
        return header
    return np.array(data, dtype=dtype, copy=True)


def _list_files(path, search_str='', end_str='', start_str=None, recursive=False):
    return [k for k in _files if k.startswith(path) and search_str in k[len(path):] and k.endswith(end_str)]


def _haversine(ra1, dec1, ra2, dec2):
    r1, d1, r2, d2 = (np.radians(np.asarray(v, dtype=float)) for v in (ra1, dec1, ra2, dec2))
    a = np.sin((d2 - d1) / 2) ** 2 + np.cos(d1) * np.cos(d2) * np.sin((r2 - r1) / 2) ** 2
    return np.degrees(2 * np.arcsin(np.sqrt(a)))


def load_reference():
    from blackbox_b200.set_bb import get_par
    from oracle import lacosmic as olac, stats as ostats
    _stub('set_zogy', mask_value=dict(MASK_VALUE), timing=False, display=False, verbose=False)
    for n in ('set_match2SSO', 'match2SSO', 'acstools', 'acstools.satdet', 'ephem', 'watchdog',
              'watchdog.observers', 'watchdog.observers.polling', 'watchdog.events', 'qc', 'ASTA', 'matplotlib',
              'matplotlib.pyplot', 'matplotlib.colors', 'fitsio', 'PIL', 'dateutil', 'dateutil.tz', 'astropy',
              'astropy.coordinates', 'astropy.time', 'astropy.units', 'astropy.visualization', 'astropy.utils',
              'astropy.utils.iers', 'astropy.io', 'astropy.io.fits', 'astropy.table', 'astropy.wcs'):
        _stub(n)
    sys.modules['watchdog.events'].FileSystemEventHandler = type('FileSystemEventHandler', (), {})
    if USE_REAL_LIBRARIES:
        # tools/pin_oracle.py --write: the two third-party algorithms as the libraries themselves
        # compute them (only possible where astropy and astroscrappy are installed)
        import importlib
        for n in [m for m in sys.modules if m == 'astropy' or m.startswith('astropy.')]:
            del sys.modules[n]
        real_stats = importlib.import_module('astropy.stats')
        real_scrappy = importlib.import_module('astroscrappy')
        sys.modules['astropy.stats'] = real_stats
        sys.modules['astroscrappy'] = real_scrappy
    else:
        _stub('astropy.stats', sigma_clipped_stats=ostats.sigma_clipped_stats, sigma_clip=ostats.sigma_clip)
        _stub('astroscrappy', detect_cosmics=olac.detect_cosmics)
    fits = types.SimpleNamespace(Header=Header)
    zogy = _stub('zogy', np=np, ndimage=ndimage, interpolate=interpolate, get_par=get_par, fits=fits, Table=_Table,
                 read_hdulist=_read_hdulist,
                 sigma_clip=ostats.sigma_clip, sigma_clipped_stats=ostats.sigma_clipped_stats,
                 log_timing_memory=lambda *a, **k: None, mem_use=lambda *a, **k: None, isfile=os.path.isfile)
    zogy.__all__ = ['np', 'ndimage', 'interpolate', 'get_par', 'fits', 'Table', 'read_hdulist', 'log_timing_memory',
                    'mem_use', 'sigma_clip', 'sigma_clipped_stats', 'isfile']
    sys.path.insert(0, os.path.join(REF, 'Settings'))
    sys.path.insert(0, REF)
    import blackbox
    blackbox.Time = Time
    blackbox.list_files = _list_files
    blackbox.haversine = _haversine
    blackbox.run_qc_check = lambda header, tel, *a, **k: None
    blackbox.get_rand_indices = lambda shape, fraction=0.2: tuple(slice(None) for _ in shape)
    blackbox.already_exists = lambda name, get_filename=False: ((name in _files, name) if get_filename
                                                               else name in _files)
    return blackbox


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def spots(a, n=8):
    flat = np.asarray(a).ravel()
    return [float(flat[i]) for i in np.linspace(0, flat.size - 1, n).astype(int)]


def sections_as_lists(secs):
    return [[[[s.start, s.stop] for s in pair] for pair in tup] for tup in secs]


def frame_case(bb, tel, seed, ysc, cosmics=True, xbin=1, variant=None):
    """gain_corr -> os_corr -> (bias) -> mask_init -> flat -> cosmics_corr -> xtalk_corr, the order of
    blackbox_reduce (blackbox.py:1479-1902), every step the reference's own function."""
    from blackbox_b200 import set_bb as my_set_bb, synth
    ref_set_bb = sys.modules['set_blackbox']
    saved = (my_set_bb.ysize_chan, ref_set_bb.ysize_chan)
    my_set_bb.ysize_chan = ref_set_bb.ysize_chan = ysc
    try:
        bb.tel = tel
        if xbin == 2:
            raw, _ = synth.make_raw(tel, seed, ysize_chan=ysc // 2, xsize_chan=660, os_rows=10, os_cols=90,
                                    nstars=300, ncosmics=60)
            out = {'tel': tel, 'seed': seed, 'ysize_chan': ysc, 'xbin': 2, 'raw_sha256': digest(raw)}
            header = Header(EXPTIME=60.0)
            data = np.array(raw, dtype='float32')
            bb.gain_corr(data, header, tel=tel)
            out['gain_sha256'] = digest(data)
            data = bb.os_corr(data, header, 'object', xbin=2, ybin=2, tel=tel)
            out['os_sha256'] = digest(data)
            out['os_spots'] = spots(data)
            out['os_header'] = {k: (v if isinstance(v, (bool, str)) else float(v)) for k, v in header.items()
                                if k.startswith(('BIASM', 'RDN', 'VFITOK', 'BIAS'))}
            return out
        raw, _ = synth.make_raw(tel, seed, nstars=400, ncosmics=150)
        if tel != 'ML1':
            raw[ysc - 50:ysc, 300:304] = 65535                   # saturated columns next to the overscan
            raw[ysc - 900:ysc - 880, 1500 * 2 + 20:1500 * 2 + 24] = 65535
        raw[40:48, 2000:2008] = 65535
        if variant == 'hos':
            synth.add_hos_contamination(raw, ysc)
        if variant == 'rings':
            synth.add_saturated_rings(raw)
        shape = (2 * ysc, 8 * my_set_bb.xsize_chan)
        mbias, mflat, bpm = synth.make_masters(tel, seed + 1, shape)
        victim, source, corr, coeffs = synth.make_xtalk(seed + 2)
        out = {'tel': tel, 'seed': seed, 'ysize_chan': ysc, 'xbin': 1, 'cosmics': bool(cosmics), 'raw_sha256': digest(raw)}
        if variant:
            out['variant'] = variant
        header = Header(EXPTIME=60.0)
        data = np.array(raw, dtype='float32')
        bb.gain_corr(data, header, tel=tel)
        out['gain_sha256'] = digest(data)
        data = bb.os_corr(data, header, 'object', tel=tel)
        out['os_sha256'] = digest(data)
        out['os_spots'] = spots(data)
        out['os_header'] = {k: (v if isinstance(v, (bool, str)) else float(v)) for k, v in header.items()
                            if k.startswith(('BIASM', 'RDN', 'VFITOK', 'BIAS')) }
        if bb.get_par(ref_set_bb.subtract_mbias, tel):
            data -= mbias
        _files.clear()
        fits_bpm = bb.get_par(ref_set_bb.bad_pixel_mask, tel).replace('bpm', 'bpm_q')
        _files[fits_bpm] = bpm
        if variant == 'rings':
            synth.add_nonfinite(data, bpm)
        data_mask, header_mask = bb.mask_init(data, header, 'q', 'object')
        out['mask_init_sha256'] = digest(data_mask)
        out['mask_counts'] = {str(b): int(((data_mask & b) != 0).sum()) for b in (1, 4, 8, 32, 64)}
        out['mask_header'] = {k: float(v) for k, v in header_mask.items()}
        data /= mflat
        if cosmics:
            data, data_mask = bb.cosmics_corr(data, header, data_mask, header_mask)
            out['cosmics_sha256'] = digest(data)
            out['cosmics_mask_sha256'] = digest(data_mask)
            out['NCOSMICS'] = float(header['NCOSMICS'])
        hm2 = Header()
        bb.mask_header(data_mask, hm2)
        out['mask_header_counts'] = {k: int(v) for k, v in hm2.items() if k.endswith('NUM')}
        path = '/tmp/_ref_xtalk_{}.txt'.format(os.getpid())
        synth.write_xtalk_file(path, victim, source, corr)
        bb.xtalk_corr(data, path, data_mask)
        os.remove(path)
        out['xtalk_sha256'] = digest(data)
        out['final_spots'] = spots(data)
        return out
    finally:
        my_set_bb.ysize_chan, ref_set_bb.ysize_chan = saved


def nonlin_case(bb, seed):
    import pickle
    from blackbox_b200 import set_bb as my_set_bb
    ref_set_bb = sys.modules['set_blackbox']
    saved = (my_set_bb.ysize_chan, my_set_bb.xsize_chan, ref_set_bb.ysize_chan, ref_set_bb.xsize_chan)
    my_set_bb.ysize_chan = ref_set_bb.ysize_chan = 48
    my_set_bb.xsize_chan = ref_set_bb.xsize_chan = 60
    try:
        bb.tel = 'BG3'
        rng = np.random.default_rng(seed)
        splines = []
        for i in range(16):
            x = np.linspace(0, 60000, 80)
            y = 2e-3 * np.sin(x / (7000.0 + 300 * i)) - 3e-7 * x + 2e-4 * rng.standard_normal(x.size)
            splines.append(interpolate.UnivariateSpline(x, y, k=3, s=x.size * 4e-8))
        data = rng.uniform(-500, 140000, size=(96, 480)).astype(np.float32)
        path = '/tmp/_ref_nonlin_{}.pkl'.format(os.getpid())
        with open(path, 'wb') as fh:
            pickle.dump(splines, fh)
        want = bb.nonlin_corr(data.copy(), path)
        os.remove(path)
        return {'seed': seed, 'input_sha256': digest(data), 'output_sha256': digest(want), 'output_spots': spots(want)}
    finally:
        my_set_bb.ysize_chan, my_set_bb.xsize_chan, ref_set_bb.ysize_chan, ref_set_bb.xsize_chan = saved


def master_case(bb, tel, imgtype, seed, ysc, date_eve='20240105', filt='q'):
    """master_prep (blackbox.py:4625-5247) on a synthetic night held in memory: the reference's
    own file selection, stack median, flat post-fix and header; write_fits is intercepted."""
    from blackbox_b200 import set_bb as my_set_bb, synth
    ref_set_bb = sys.modules['set_blackbox']
    saved = (my_set_bb.ysize_chan, ref_set_bb.ysize_chan)
    my_set_bb.ysize_chan = ref_set_bb.ysize_chan = ysc
    written = {}

    def write_fits(fits_out, data, header, **k):
        written['name'], written['data'], written['header'] = fits_out, np.array(data), dict(header)
        return fits_out

    bb.write_fits = write_fits
    try:
        bb.tel = tel
        shape = (2 * ysc, 8 * my_set_bb.xsize_chan)
        red_dir = bb.get_par(ref_set_bb.red_dir, tel)
        master_dir = bb.get_par(ref_set_bb.master_dir, tel)
        _files.clear()
        for name, frame, hdr in synth.make_cal_night(tel, imgtype, seed, shape, date_eve, filt):
            _files['{}/{}'.format(red_dir, name)] = (frame, Header(hdr))
        if imgtype == 'flat':
            bpm = synth.make_masters(tel, seed + 1, shape)[2]
            _files[bb.get_par(ref_set_bb.bad_pixel_mask, tel).replace('bpm', 'bpm_' + filt)] = bpm
        tail = '_' + filt if imgtype == 'flat' else ''
        fits_master = '{}/{}/{}/{}/{}/{}_{}_{}{}.fits'.format(master_dir, date_eve[0:4], date_eve[4:6], date_eve[6:8],
                                                          imgtype, tel, imgtype, date_eve, tail)
        got = bb.master_prep(fits_master, shape, True, pick_alt=False, tel=tel, proc_mode=None)
        assert got == fits_master and written['name'] == fits_master, (got, written.get('name'))
        skip = ('DATEFILE', 'MFMED', 'MFSTD', 'MBMEAN', 'MBRDN')
        hdr = {k: (v if isinstance(v, (bool, str)) else (int(v) if isinstance(v, (int, np.integer)) else float(v)))
               for k, v in written['header'].items()
               if k not in skip and not k.startswith(('MBIASM', 'MBRDN'))}
        return {'tel': tel, 'imgtype': imgtype, 'seed': seed, 'ysize_chan': ysc, 'date_eve': date_eve, 'filt': filt,
                'master_sha256': digest(written['data'].astype(np.float32)), 'master_spots': spots(written['data']),
                'header': hdr}
    finally:
        my_set_bb.ysize_chan, ref_set_bb.ysize_chan = saved
        _files.clear()


def main():
    bb = load_reference()
    out = {'reference_version': bb.__version__, 'numpy': np.__version__,
           'third_party': 'astropy / astroscrappy as installed' if USE_REAL_LIBRARIES else 'oracle restatements (unpinned)',
           'sections': {}, 'frames': [], 'nonlin': [], 'masters': []}
    for shape, xb in (((10600, 12000), 1), ((5300, 6000), 2), ((10560, 10560), 1)):
        out['sections']['{}x{}_bin{}'.format(shape[0], shape[1], xb)] = sections_as_lists(
            bb.define_sections(shape, xbin=xb, ybin=xb, tel='BG3'))
    out['frames'].append(frame_case(bb, 'ML1', 1001, 200))
    out['frames'].append(frame_case(bb, 'BG3', 4001, 2640))
    out['frames'].append(frame_case(bb, 'BG2', 4002, 5280, cosmics=False))     # full size: channel-9 split fit
    out['frames'].append(frame_case(bb, 'ML1', 5001, 400, xbin=2))             # 2x2 binned frame
    out['frames'].append(frame_case(bb, 'ML1', 1002, 200, variant='hos'))      # charge in the horizontal overscan
    out['frames'].append(frame_case(bb, 'ML1', 1003, 200, variant='rings'))    # saturated rings: hole filling
    out['frames'][-1]['cpu_only'] = True       # added after the round's GPU time was spent: oracle test only
    out['nonlin'].append(nonlin_case(bb, 11))
    out['masters'] = [master_case(bb, 'ML1', 'bias', 7001, 40), master_case(bb, 'BG3', 'flat', 7101, 2000)]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_golden.json')
    with open(path, 'w') as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print('wrote', path)
    return out


if __name__ == '__main__':
    main()
