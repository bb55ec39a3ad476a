"""``master_prep``: the reference's master-frame entry point over FITS files, with the combine on
the GPU (blackbox.py:4625-5247; SURVEY.md 8b lists its signature as part of the boundary).

What is mirrored, step by step:

  * the file name grammar ``[tel]_[imgtype]_[date_eve][_filt].fits``           blackbox.py:4646-4659
  * an existing, not red-flagged master is returned as is                      :4663-4676
  * the calibration frames of ``date_eve`` +/- ``set_bb.cal_window`` days are listed from
    ``[red_dir]/yyyy/mm/dd/[imgtype]/[tel]_20*[filt].fits*``                    :4698-4730
  * red-flagged frames, MeerLICHT evening flats of 2019-07..2020-03 and BlackGEM evening
    flats are dropped                                                          :4746-4794
  * fewer than 5 frames, a red-flagged master or ``create_master=False``: the nearest good
    master (yesterday's, else the nearest in the previous/current/next month)  :4802-4847, 5302-5395
  * at most ``set_bb.ncal_max`` frames nearest to midnight of the evening date :4855-4868
  * no new master if all of them are older than 12 hours                       :4876-4884
  * the stack median (flats divided by MEDSEC or the median over ``flat_norm_sec``) and the
    flat post-fix: ``reduce.master_combine`` -> ``bbx_stack_median``           :4908-4984, 5063-5073
  * the master header: names of the frames, N[TYPE], [TYPE]-WIN, and for flats STATSEC,
    MFMEDSEC, MFSTDSEC, N-OFFSET, OFF-MEAN, FLATDITH and the channel factors GAINCF1..16
                                                                               :4945-5161

Left out on purpose: the header statistics drawn from a RANDOM subsample (MFMED / MFSTD, MBMEAN /
MBRDN, MBIASM / MBRDN per channel; ``get_rand_indices`` of the absent zogy module, not
reproducible), the QC flags (``run_qc_check``, qc.py: policy, out of scope), fpack and the jpg.
Files are read with ``reduce.read_fits_image``: plain primary HDUs or fpacked ``.fits.fz`` (what the
reference's folders hold; Rice tiles decoded and float images un-quantised on the GPU), the
bad-pixel mask likewise; the master is written as an uncompressed float32 primary HDU.  An input
this reader cannot unpack ends, as any unusable input does in the reference, with the nearest
existing master and a logged error.

The individual frames go to the GPU one by one (pinned big-endian bytes, decoded there); the
whole stack stays resident (15 x 446 MB for a full-size master flat)."""
import datetime
import glob
import logging
import os
import time

import numpy as np
import torch

from . import fitsio, set_bb
from . import reduce as R
from .geometry import define_sections
from .set_bb import get_par

log = logging.getLogger(__name__)

NIGHT_WAIT_S = 60                    # blackbox.py:4690-4695 (old "chopper" night mode)
_MJD0 = datetime.datetime(1858, 11, 17)


# -------------------------------------------------------------------------------------------
# small stand-ins for astropy.time / zogy helpers
# -------------------------------------------------------------------------------------------
def date2mjd(date_str, time_str=None):
    """blackbox.py:5416-5440 (UTC scale, no leap-second arithmetic is involved in a UTC MJD)."""
    if '-' not in date_str:
        date_str = '{}-{}-{}'.format(date_str[0:4], date_str[4:6], date_str[6:8])
    fmt = '%Y-%m-%d'
    if time_str is not None:
        if ':' not in time_str:
            time_str = '{}:{}:{}'.format(time_str[0:2], time_str[2:4], time_str[4:])
        date_str = '{} {}'.format(date_str, time_str)
        fmt += ' %H:%M' if time_str.count(':') == 1 else ' %H:%M:%S'
        if '.' in time_str:
            fmt += '.%f'
    t = datetime.datetime.strptime(date_str, fmt)
    delta = t - _MJD0
    return delta.days + (delta.seconds + delta.microseconds * 1e-6) / 86400.0


def mjd2date(mjd):
    """-> 'yyyy/mm/dd' of an MJD (``Time(mjd, format='mjd').isot.split('T')[0].replace('-','/')``)."""
    return (_MJD0 + datetime.timedelta(days=float(mjd))).strftime('%Y/%m/%d')


already_exists = fitsio.already_exists          # blackbox.py:787-807


def list_files(path, search_str='', end_str='', start_str=None, recursive=False):
    """The file-system branch of zogy's list_files (its copy: blackbox_slurm_google.py:1337-1372)."""
    if os.path.isdir(path) and not path.endswith('/'):
        path += '/'
    folder, prefix = os.path.split(path)
    if not folder:
        folder, prefix = prefix, ''
    if prefix == '' and start_str is not None:
        prefix = start_str
    if recursive:
        files = glob.glob('{}/**/{}*{}*{}'.format(folder, prefix, search_str, end_str), recursive=True)
        if path in files:
            files.remove(path)
        return files
    return glob.glob('{}/{}*{}*{}'.format(folder, prefix, search_str, end_str))


def read_header(path):
    """Header of the image in ``path`` as a plain dict key -> value: the primary header, merged
    with the first extension's for files whose primary HDU is empty (fpacked files keep the image
    keywords, QC-FLAG among them, in the extension)."""
    with open(path, 'rb') as fh:
        hdr, _ = fitsio.read_header(fh)
        if hdr.get('NAXIS', (0,))[0] == 0:
            try:
                ext, _ = fitsio.read_header(fh)
                for k, v in ext.items():
                    hdr.setdefault(k, v)
            except fitsio.FitsError:
                pass
    return {k: v[0] for k, v in hdr.items()}, {k: v[1] for k, v in hdr.items()}


def qc_flagged(path, flag='red'):
    """blackbox.py:5403-5411."""
    return read_header(path)[0].get('QC-FLAG') == flag


def haversine(ra1, dec1, ra2, dec2):
    """Great-circle distance in degrees (zogy's haversine, used for the dithering check
    blackbox.py:5034-5037)."""
    r1, d1, r2, d2 = (np.radians(np.asarray(v, dtype=float)) for v in (ra1, dec1, ra2, dec2))
    a = np.sin((d2 - d1) / 2) ** 2 + np.cos(d1) * np.cos(d2) * np.sin((r2 - r1) / 2) ** 2
    return np.degrees(2 * np.arcsin(np.sqrt(a)))


# -------------------------------------------------------------------------------------------
def _name_parts(fits_master, tel):
    filename = os.path.split(fits_master)[1]
    stem = filename.split('.fits')[0]
    imgtype, date_eve = stem.split('{}_'.format(tel))[-1].split('_')[0:2]
    filt = stem.split('_')[-1] if imgtype == 'flat' else None
    return imgtype, date_eve, filt


def delta_one_month(date_eve, dmonth):
    """'yyyy/mm/' of the previous / current / next month (blackbox.py:5251-5290)."""
    date_eve = ''.join(c for c in date_eve if c.isdigit())
    if dmonth == 0:
        mjd_noon = date2mjd(date_eve, time_str='12:00')
    elif dmonth == -1:
        mjd_noon = date2mjd(date_eve, time_str='12:00') - (int(date_eve[6:8]) + 1)
    elif dmonth == 1:
        year, month = int(date_eve[0:4]), int(date_eve[4:6])
        year, month = (year + 1, 1) if month == 12 else (year, month + 1)
        mjd_noon = date2mjd('{}{:02}{:02}'.format(year, month, 1), time_str='12:00')
    else:
        raise ValueError('maximum [dmonth] in [delta_one_month] is 1')
    return mjd2date(mjd_noon)[0:8]


def get_nearest_master(date_eve, imgtype, fits_master, filt=None, tel=None):
    """blackbox.py:5294-5395: yesterday's master if it is there and not red-flagged, else the
    nearest (in evening date) unflagged master of the previous, current and next month."""
    dash = '{}-{}-{}'.format(date_eve[0:4], date_eve[4:6], date_eve[6:8])
    slash = dash.replace('-', '/')
    yest = datetime.datetime.strptime(dash, '%Y-%m-%d') - datetime.timedelta(days=1)
    yest_dash = yest.strftime('%Y-%m-%d')
    fits_yest = (fits_master.replace(date_eve, yest_dash.replace('-', ''))
                 .replace(slash, yest_dash.replace('-', '/')))
    present, name = already_exists(fits_yest, get_filename=True)
    if present and not qc_flagged(name):
        return name
    master_dir = get_par(set_bb.master_dir, tel)
    files = []
    for n_month in (-1, 0, 1):
        path_tmp = '{}/{}'.format(master_dir, delta_one_month(date_eve, n_month))
        end_str = '{}.fits*'.format(filt) if imgtype == 'flat' else '.fits*'
        files.extend(list_files(path_tmp, start_str='{}_{}_'.format(tel, imgtype), end_str=end_str,
                                recursive=True))
    files = sorted(set(files))
    if not files:
        return None
    mjds = np.array([date2mjd(''.join(f.split('/')[-5:-2])) for f in files])
    for i_near in np.argsort(abs(mjds - date2mjd(date_eve))):
        if not qc_flagged(files[i_near]):
            return files[i_near]
    return None


def flat_channel_factors(master, tel=None):
    """GAINCF1..16 (blackbox.py:5081-5153): factors that would level the channels of a master
    flat -- vertically from the 200 rows either side of the read-out boundary, horizontally from
    2000 x 200-pixel strips either side of each channel boundary -- normalised to a mean of one.
    ``master``: float32 CUDA tensor (not modified).  The dtypes follow numpy's: the vertical
    step divides float32 data by a float64 median (division in double, rounded to float32),
    the horizontal ratios are float32."""
    data_sec_red = define_sections(tuple(master.shape), tel=tel)[4]
    nchans = len(data_sec_red)
    corr = master.clone()
    med = np.zeros(nchans)
    nrows = 200
    for i, sec in enumerate(data_sec_red):
        chan = corr[sec]
        med[i] = R.exact_median(chan[-nrows:, :] if i < 8 else chan[0:nrows, :])
        corr[sec] = (chan.double() / float(med[i])).float()
    factor = 1.0 / med
    ysize, xsize = master.shape
    ny, nx = get_par(set_bb.ny, tel), get_par(set_bb.nx, tel)
    dy, dx = ysize // ny, xsize // nx
    nrows, ncols = 2000, 200
    for i in range(1, nx):
        y0, x0 = dy, i * dx
        rows = _np_slice(y0 - nrows, y0 + nrows, ysize)
        stat1 = corr[rows, _np_slice(x0 - ncols, x0, xsize)]
        stat2 = corr[rows, _np_slice(x0, x0 + ncols, xsize)]
        ratio = np.float32(R.exact_median(stat1)) / np.float32(R.exact_median(stat2))
        for sec in (data_sec_red[i], data_sec_red[i + nx]):
            corr[sec] *= float(ratio)
        factor[i] *= ratio
        factor[i + nx] *= ratio
    return factor / np.mean(factor)


def _np_slice(start, stop, size):
    """numpy's reading of ``a[start:stop]`` for possibly negative bounds, as a torch-safe slice."""
    return slice(*slice(start, stop).indices(size)[:2])


# -------------------------------------------------------------------------------------------
def master_prep(fits_master, data_shape, create_master, pick_alt=True, tel=None, proc_mode=None):
    """Create the master calibration frame ``fits_master`` of shape ``data_shape`` unless a good
    one exists, or pick a nearby one (see the module text).  -> path of the master, or None."""
    imgtype, date_eve, filt = _name_parts(fits_master, tel)
    master_present, fits_master = already_exists(fits_master, get_filename=True)
    master_ok = True
    if master_present:
        log.info('%s master already on disk: %s', imgtype, fits_master)
        if qc_flagged(fits_master):
            master_ok = False
            log.warning('%s master %s is red-flagged', imgtype, fits_master)
    if master_present and master_ok:
        return fits_master

    if proc_mode == 'night' and not fits_master.startswith('gs://') and NIGHT_WAIT_S > 0:
        log.warning('night mode: giving the calibration frames %d s to land on disk', NIGHT_WAIT_S)
        time.sleep(NIGHT_WAIT_S)

    nwindow = int(get_par(set_bb.cal_window, tel)[imgtype])
    red_dir = get_par(set_bb.red_dir, tel)
    file_list = []
    for n_day in range(-nwindow, nwindow + 1):
        mjd_noon = date2mjd(date_eve, time_str='12:00') + n_day
        path_tmp = '{}/{}/{}/{}_20'.format(red_dir, mjd2date(mjd_noon), imgtype, tel)
        search_str = '{}.fits'.format(filt) if imgtype == 'flat' else '.fits'
        file_list.extend(list_files(path_tmp, search_str=search_str))
    file_list = sorted(file_list)
    nfiles = len(file_list)

    if create_master:
        mjd_obs = np.zeros(nfiles)
        keep = np.ones(nfiles, dtype=bool)
        mjd_avoid = (date2mjd('2019-07-01', '12:00:00'), date2mjd('2020-03-01', '12:00:00'))
        for i, name in enumerate(file_list):
            hdr = read_header(name)[0]
            if hdr.get('QC-FLAG') == 'red':
                keep[i] = False
            if 'MJD-OBS' in hdr:
                mjd_obs[i] = hdr['MJD-OBS']
            if tel == 'ML1' and mjd_obs[i] % 1 > 0.5 and mjd_avoid[0] < mjd_obs[i] < mjd_avoid[1]:
                keep[i] = False
            if imgtype == 'flat' and get_par(set_bb.flat_reject_eve, tel) and \
                    (mjd_obs[i] % 1 > 0.5 or mjd_obs[i] % 1 < 0.1):
                log.warning('evening flat left out: %s', name)
                keep[i] = False
        file_list = np.array(file_list)[keep]
        mjd_obs = mjd_obs[keep]
        nfiles = len(file_list)

    msg = 'flat in filter {}'.format(filt) if imgtype == 'flat' else imgtype
    wanted = (imgtype == 'bias' and get_par(set_bb.subtract_mbias, tel)) or imgtype == 'flat'
    if nfiles < 5 or not master_ok or not create_master:
        if not (pick_alt or not create_master):
            if master_ok:
                log.warning('fewer than 5 usable frames for a master %s within %d days of %s', msg, nwindow,
                            date_eve)
            return None
        near = get_nearest_master(date_eve, imgtype, fits_master, filt=filt, tel=tel)
        if near is None:
            if wanted:
                log.error('no nearby master %s to fall back on', msg)
            return None
        if wanted:
            log.warning('evening date %s: falling back on master %s', date_eve, near)
        return near

    nmax = int(get_par(set_bb.ncal_max, tel)[imgtype])
    delta = mjd_obs - date2mjd(date_eve, time_str='23:59')
    order = np.argsort(np.abs(delta))
    file_list = file_list[order][0:nmax]
    delta = delta[order][0:nmax]
    nfiles_orig, nfiles = nfiles, len(file_list)
    if np.amin(np.abs(delta)) > 0.5 and np.all(delta < 0):
        log.warning('the %d frames nearest to midnight of %s all predate it by more than 12 hours; %s would '
                    'repeat an older master and is not made', nmax, date_eve, fits_master)
        return None
    log.info('%s master %s of %s from:\n%s', tel, msg, date_eve, file_list)
    if nfiles_orig > nmax:
        log.warning('%d %s frames found, ncal_max is %d: keeping the ones nearest to midnight of %s', nfiles_orig,
                    imgtype, nmax, date_eve)

    try:
        master, header = combine_files(list(file_list), tuple(data_shape), imgtype, filt, nwindow, tel)
    except (fitsio.FitsError, NotImplementedError) as exc:
        # an input this reader cannot unpack (a tile kept in a gzip fall-back column, say): as for
        # any other unusable input the reference ends up with a nearby master, not with a raise
        log.error('master %s of %s not made, unreadable input: %s', msg, date_eve, exc)
        if not pick_alt:
            return None
        return get_nearest_master(date_eve, imgtype, fits_master, filt=filt, tel=tel)
    return write_master(fits_master, master, header)


def write_master(fits_master, master, header):
    """The master as a float32 primary HDU (big-endian bytes made on the GPU); the reference's
    write_fits(..., master=True) (blackbox.py:5238-5240, 7653-7675) minus fpack and the jpg."""
    now = datetime.datetime.now(datetime.timezone.utc).replace(tzinfo=None)
    header['DATEFILE'] = (now.isoformat(timespec='milliseconds'), 'UTC date of writing file')
    os.makedirs(os.path.dirname(os.path.abspath(fits_master)), exist_ok=True)
    be, bitpix = R.fits_encode(master)
    fitsio.write_primary(fits_master, be.cpu(), header, be_bytes=True, shape=tuple(master.shape), bitpix=bitpix)
    return fits_master


def combine_files(file_list, data_shape, imgtype, filt, nwindow, tel):
    """Read the frames onto the GPU, combine them, and build the master header
    (blackbox.py:4905-5161).  -> (float32 CUDA tensor, header dict key -> (value, comment))."""
    nfiles = len(file_list)
    dev = torch.device('cuda', torch.cuda.current_device())
    frames, medsec = [], []
    header = {}
    ra, dec = [], []
    up = imgtype.upper()
    for i, name in enumerate(file_list):
        # plain or fpacked (the reference's red_dir holds .fits.fz: Rice-coded, float images
        # quantised; blackbox.py:4698-4730 lists '*.fits*'): decoded on the GPU either way
        val, frame = R.read_fits_image(name, dtype=torch.float32)
        com = read_header(name)[1]
        if tuple(frame.shape) != tuple(data_shape):
            raise ValueError('{}: shape {} instead of {}'.format(name, tuple(frame.shape), tuple(data_shape)))
        frames.append(frame)
        if imgtype == 'flat':
            medsec.append(val.get('MEDSEC'))
            if 'RA' in val and 'DEC' in val:
                ra.append(val['RA'])
                dec.append(val['DEC'])
        if i == 0:
            for key in ('IMAGETYP', 'DATE-OBS', 'FILTER', 'RA', 'DEC', 'XBINNING', 'YBINNING', 'MJD-OBS',
                        'AIRMASS', 'ORIGIN', 'TELESCOP', 'PYTHON-V', 'BB-V'):
                if key in val:
                    header[key] = (val[key], com[key])
        comment = 'name reduced flat' if imgtype == 'flat' else 'name gain/os-corrected {} frame'.format(imgtype)
        header['{}{}'.format(up, i + 1)] = (name.split('/')[-1].split('.fits')[0], '{} {}'.format(comment, i + 1))
        if 'ORIGFILE' in val:
            header['{}OR{}'.format(up, i + 1)] = (val['ORIGFILE'], 'name original {} {}'.format(imgtype, i + 1))
        if i == nfiles - 1:
            for key in ('DATE-END', 'MJD-END'):
                if key in val:
                    header[key] = (val[key], com[key])

    bpm = None
    if imgtype == 'flat':
        fits_bpm = get_par(set_bb.bad_pixel_mask, tel).replace('bpm', 'bpm_{}'.format(filt))
        present, fits_bpm = already_exists(fits_bpm, get_filename=True)
        if present:
            _, bpm = R.read_fits_image(fits_bpm)               # plain or fpacked (set_blackbox.py:187-193: .fits.fz)
            if bpm.dtype != torch.uint8:
                raise fitsio.FitsError('{}: bad-pixel mask is not an 8-bit image ({})'.format(fits_bpm, bpm.dtype))
    master, scales = R.master_combine(frames, imgtype=imgtype, medsec=medsec if imgtype == 'flat' else None,
                                      bpm=bpm, tel=tel)
    if imgtype == 'flat':
        # the reference takes MFMEDSEC / MFSTDSEC BEFORE the edge / non-positive post-fix
        # (blackbox.py:5006-5013 vs 5063-5073): recombine the rows of the statistics section
        # without the fix (the fix is fused into the combine kernel)
        sec = get_par(set_bb.flat_norm_sec, tel)
        r0, r1 = sec[0].indices(data_shape[0])[:2]
        if r1 > r0:
            rows, _ = R.master_combine([f[r0:r1] for f in frames], imgtype='flat', medsec=scales, bpm=None, tel=tel)
            stat_sec = rows[:, sec[1]]
        else:
            stat_sec = master[sec]
    del frames

    header['N{}'.format(up)] = (nfiles, 'number of {} frames combined'.format(imgtype.lower()))
    header['{}-WIN'.format(up)] = (nwindow, '[days] input time window to include {} frames'.format(imgtype.lower()))
    if imgtype == 'flat':
        header['STATSEC'] = ('[{}:{},{}:{}]'.format(sec[0].start + 1, sec[0].stop + 1, sec[1].start + 1,
                                                     sec[1].stop + 1),
                             'pre-defined statistics section [y1:y2,x1:x2]')
        header['MFMEDSEC'] = (R.exact_median(stat_sec), 'median master flat over STATSEC')
        header['MFSTDSEC'] = (float(torch.std(stat_sec.double(), unbiased=False).float()),
                              'sigma (STD) master flat over STATSEC')
        noffset, offset_mean = 0, 0
        if len(ra) > 0 and len(dec) > 0:
            ra, dec = np.array(ra), np.array(dec)
            offset = 3600. * haversine(ra, dec, np.roll(ra, 1), np.roll(dec, 1))
            off = offset >= 5
            noffset = int(np.sum(off))
            if noffset > 0:
                offset_mean = float(np.mean(offset[off]))
        header['N-OFFSET'] = (noffset, 'number of flats with offsets > 5 arcsec')
        header['OFF-MEAN'] = (offset_mean, '[arcsec] mean dithering offset')
        header['FLATDITH'] = (bool(float(noffset) / nfiles >= 0.66), 'majority of flats were dithered')
        factor = flat_channel_factors(master, tel=tel)
        for i in range(len(factor)):
            header['GAINCF{}'.format(i + 1)] = (float(factor[i]), 'channel {} gain correction factor'.format(i + 1))
    return master, header
