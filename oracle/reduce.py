"""CPU ORACLE (test infrastructure only): the reference's per-frame reduction steps.

numpy/scipy restatement of blackbox.py's hot-path functions, same call surface, same
arithmetic and dtypes (numpy >= 2 promotion rules, which is what is installed here):

    gain_corr     blackbox.py:7442-7465
    os_corr       blackbox.py:6407-6879
    mask_init     blackbox.py:4375-4579   (+ fill_sat_holes 4584-4596)
    cosmics_corr  blackbox.py:4259-4370   (detect_cosmics -> oracle.lacosmic)
    xtalk_corr    blackbox.py:7138-7258
    master_median blackbox.py:4908-4984, 5063-5073 (arithmetic core of master_prep)
    reduce_frame  blackbox.py:1479-1902   (the order blackbox_reduce calls them in)

numpy, scipy.ndimage and scipy.interpolate are called exactly as the reference calls them;
astropy's sigma clipping and astroscrappy come from oracle.stats / oracle.lacosmic (restated,
UNPINNED).  File and FITS handling is replaced by in-memory arguments (BPM array, coefficient
matrix, master arrays).  Every function optionally returns its intermediates (``diag``) so
single kernels can be checked.

Known, deliberate differences from the reference (edge behaviour only):
  * the reference turns every warning into an exception inside os_corr
    (warnings.filterwarnings('error'), blackbox.py:6432); here only the spline/polyfit calls
    run under that filter, so numpy "empty slice" RuntimeWarnings (e.g. a horizontal-overscan
    column with exactly one surviving value) do not abort the overscan correction.
"""
import warnings

import numpy as np
from scipy import interpolate, ndimage

from blackbox_b200 import set_bb
from blackbox_b200.geometry import define_sections
from blackbox_b200.set_bb import get_par

from . import lacosmic
from .stats import sigma_clip, sigma_clipped_stats

F32 = np.float32


# -------------------------------------------------------------------------------------------
def gain_corr(data, header, tel=None):
    """In place ``data[chan] *= gain[chan]`` (float32 multiply); blackbox.py:7442-7465."""
    gain = get_par(set_bb.gain, tel)
    chan_sec = define_sections(data.shape, tel=tel)[0]
    for i, sec in enumerate(chan_sec):
        data[sec] *= gain[i]
        header['GAIN{}'.format(i + 1)] = gain[i]


# -------------------------------------------------------------------------------------------
def hos_column_stats(data_hos, mask_hos, nsigma=2.5):
    """Column-wise clipped mean / std(ddof=1) / count of the horizontal-overscan strip,
    float32 accumulation in row order (np.nanmean / np.nanstd on the float32 MaskedArray
    that sigma_clip returns; blackbox.py:6649-6662)."""
    clipped = sigma_clip(np.ma.masked_array(data_hos, mask=mask_hos), axis=0,
                         cenfunc='mean', sigma=nsigma)
    keep = ~np.ma.getmaskarray(clipped)
    n = keep.sum(axis=0)
    with np.errstate(invalid='ignore', divide='ignore'):
        tot = np.add.reduce(np.where(keep, data_hos, F32(0)), axis=0, dtype=F32)
        mean = (tot.astype(np.float64) / n).astype(F32)
        dev = np.where(keep, data_hos - mean[None, :], F32(0)).astype(F32)
        var = np.add.reduce(dev * dev, axis=0, dtype=F32)
        var = (var.astype(np.float64) / (n - 1)).astype(F32)
        std = np.sqrt(var)
    bad = n - 1 <= 0
    std[bad] = np.nan
    mean[n == 0] = np.nan
    return mean, std, n


def _ml1_hos_mask(data_hos, data_limit):
    """blackbox.py:6586-6614"""
    mask_hos = data_hos > data_limit
    mask_x = np.sum(mask_hos, axis=0) > 0.5 * mask_hos.shape[0]
    mask_x_open = ndimage.binary_opening(mask_x, structure=np.ones(2))
    mask_hos[:, np.logical_xor(mask_x, mask_x_open)] = False
    return ndimage.binary_dilation(mask_hos, structure=np.ones((3, 3), dtype=bool),
                                   iterations=2)


def _bg_satcol(data, data_sec_i, i_chan, nrows, sat_e, tel):
    """blackbox.py:6624-6640"""
    lim = get_par(set_bb.hos_sat_ypix_lim, tel)
    if i_chan >= 8:
        r1, r2 = range(0, lim[0]), range(0, lim[1])
    else:
        r1, r2 = range(nrows - lim[0], nrows), range(nrows - lim[1], nrows)
    sec = data[data_sec_i]
    thr = 0.9 * sat_e
    satcol = np.sum(sec[r1, :] >= thr, axis=0) >= 3
    satcol |= np.sum(sec[r2, :] >= thr, axis=0) >= 10
    return satcol


def running_median3(y):
    """3-point running median of y[3:], windows clipped to [3, n), built from the
    un-smoothed values (blackbox.py:6703-6708)."""
    y = np.array(y, copy=True)
    n = len(y)
    y[3:] = [np.median(y[max(k - 1, 3):min(k + 2, n)]) for k in range(3, n)]
    return y


def hos_fit(mean_hos, std_hos, nvalues, satcol, tel, i_chan, diag=None):
    """Columns -> overscan vector to subtract (float64 [ncols]); blackbox.py:6662-6814.
    ``satcol`` is None for ML1."""
    ncols = len(mean_hos)
    mask_valid = nvalues > 1
    xcol = np.arange(ncols) + 1
    err_hos = np.zeros(ncols, dtype=F32)
    err_hos[mask_valid] = std_hos[mask_valid] / np.sqrt(nvalues[mask_valid])
    weights = np.zeros(ncols, dtype=F32)
    nz = err_hos != 0
    weights[nz] = 1 / err_hos[nz]
    if np.all(mask_valid[0:3]):
        weights[0:3] = 0

    idx_switch, overlap = 150, 30
    idx_fit = np.arange(min(idx_switch + overlap, ncols))
    npoints = int(np.sum(mask_valid[idx_fit] & nz[idx_fit]))
    m = mask_valid
    y2fit = running_median3(mean_hos[idx_fit][m[idx_fit]])
    xs = xcol[idx_fit][m[idx_fit]]
    ws = weights[idx_fit][m[idx_fit]]
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        try:
            splfit = interpolate.UnivariateSpline(xs, y2fit, w=ws, k=2, s=npoints)
            retried = False
        except UserWarning:
            splfit = interpolate.UnivariateSpline(xs, y2fit, w=ws, k=3, s=1.5 * npoints)
            retried = True

    mask_valid_poly = mask_valid.copy()
    mask_valid_poly[0:idx_switch - overlap] = False
    mean_hos_poly = mean_hos[mask_valid_poly]
    mean, _, stddev = sigma_clipped_stats(mean_hos_poly, sigma=5, cenfunc='mean')
    if stddev == 0:
        keep = np.ones(len(mean_hos_poly), dtype=bool)
    else:
        keep = np.abs(mean_hos_poly - mean) / stddev <= 5
    mask_valid_poly[mask_valid_poly] = keep

    def fit3(mask_fit, deg):
        fit = None
        for _ in range(3):
            with warnings.catch_warnings():
                warnings.simplefilter('error')
                p = np.polyfit(xcol[mask_fit], mean_hos[mask_fit], deg)
            fit = np.polyval(p, xcol)
            with np.errstate(invalid='ignore'):
                mask_fit &= np.abs(fit - mean_hos) <= 3 * err_hos
        return fit

    if not (tel == 'BG2' and i_chan == 8):
        oscan = fit3(mask_valid_poly, 7)
    else:
        idx_split = 654
        mf = mask_valid_poly.copy()
        mf[idx_split:] = False
        fit1 = fit3(mf, 5)
        mf = mask_valid_poly.copy()
        mf[:idx_split] = False
        fit2 = fit3(mf, 5)
        oscan = fit1
        oscan[idx_split:] = fit2[idx_split:]

    spline_vals = splfit(xcol[0:idx_switch])
    oscan[0:idx_switch] = spline_vals
    first = np.arange(ncols) < 3
    sel = first & mask_valid
    oscan[sel] = mean_hos[sel]
    use_mean = mask_valid.copy()
    if tel[0:2] == 'BG':
        use_mean &= ~satcol
    use_mean[idx_switch:] = False
    oscan[use_mean] = mean_hos[use_mean]
    if diag is not None:
        need = np.zeros(ncols, dtype=bool)
        need[:idx_switch] = True
        need &= ~(use_mean | sel)
        diag.update(err_hos=err_hos, weights=weights, spline=spline_vals,
                    spline_retried=retried, spline_needed=need, use_mean=use_mean)
    return oscan


def os_corr(data, header, imgtype, xbin=1, ybin=1, data_limit=2000, tel=None, diag=None):
    """Overscan correction; returns the cropped float32 frame, mutates ``data`` and
    ``header``; blackbox.py:6407-6879.  ``diag`` (dict) receives per-channel intermediates."""
    chan_sec, data_sec, os_sec_hori, os_sec_vert, data_sec_red = define_sections(
        data.shape, xbin=xbin, ybin=ybin, tel=tel)
    ncols = get_par(set_bb.xsize_chan, tel) // xbin
    nrows = get_par(set_bb.ysize_chan, tel) // ybin
    ny, nx = get_par(set_bb.ny, tel), get_par(set_bb.nx, tel)
    data_out = np.zeros((nrows * ny, ncols * nx), dtype=F32)
    nchans = len(data_sec)
    mean_vos = np.zeros(nchans)
    std_vos = np.zeros(nchans)
    vos_poldeg = get_par(set_bb.voscan_poldeg, tel)
    nrows_chan = data[chan_sec[0]].shape[0]
    y_vos = np.arange(nrows_chan)
    nrows_overlap = nrows_chan - nrows
    sat_e = np.array(get_par(set_bb.satlevel, tel)) * np.array(get_par(set_bb.gain, tel))
    chans = []

    for i in range(nchans):
        d = {}
        # vertical overscan: clipped mean per row, low-order polynomial along y
        mean_vos_col = sigma_clipped_stats(data[os_sec_vert[i]], axis=1, mask_value=0,
                                           cenfunc='mean')[0]
        polyfit_ok = True
        mean, _, stddev = sigma_clipped_stats(mean_vos_col, sigma=5, cenfunc='mean')
        if stddev == 0:
            mask_fit = np.ones(nrows_chan, dtype=bool)
        else:
            with np.errstate(invalid='ignore'):
                mask_fit = np.abs(mean_vos_col - mean) / stddev <= 5
        if i < 8:
            mask_fit[nrows:] = False
        else:
            mask_fit[:nrows_overlap] = False
        with warnings.catch_warnings():
            warnings.simplefilter('error')
            p = np.polyfit(y_vos[mask_fit], mean_vos_col[mask_fit], vos_poldeg)
        for nc in range(len(p)):
            c = p[::-1][nc]
            header['BIAS{}A{}'.format(i + 1, nc)] = c if np.isfinite(c) else 'None'
        fit_vos_col = np.polyval(p, y_vos)
        if not np.all(np.isfinite(fit_vos_col)):
            polyfit_ok = False
        header['VFITOK{}'.format(i + 1)] = polyfit_ok
        if polyfit_ok:
            mean_vos[i] = np.mean(fit_vos_col)
            data[chan_sec[i]] -= fit_vos_col.reshape(nrows_chan, 1)
        else:
            mean_vos[i] = np.nanmedian(mean_vos_col)
            data[chan_sec[i]] -= mean_vos[i]

        # level offset between vertical and horizontal overscan
        dlevel = sigma_clipped_stats(data[os_sec_hori[i]][:, ncols - 300:ncols],
                                     cenfunc='mean')[0]
        data[os_sec_hori[i]] -= dlevel
        std_vos[i] = sigma_clipped_stats(data[os_sec_vert[i]], mask_value=0,
                                         cenfunc='mean')[2]

        # horizontal overscan
        data_hos = data[os_sec_hori[i]][:, :ncols]
        satcol = None
        if tel == 'ML1':
            mask_hos = _ml1_hos_mask(data_hos, data_limit)
        else:
            satcol = _bg_satcol(data, data_sec[i], i, nrows, sat_e[i], tel)
            mask_hos = np.zeros(data_hos.shape, dtype=bool)
            mask_hos[:] |= satcol
        mean_hos, std_hos, nvalues = hos_column_stats(data_hos, mask_hos)
        oscan = hos_fit(mean_hos, std_hos, nvalues, satcol, tel, i, diag=d)

        data[data_sec[i]] -= oscan
        data_out[data_sec_red[i]] = data[data_sec[i]]
        d.update(mean_vos_col=mean_vos_col, mask_fit=mask_fit, p=p, fit_vos_col=fit_vos_col,
                 polyfit_ok=polyfit_ok, dlevel=dlevel, mask_hos=mask_hos, satcol=satcol,
                 mean_hos=mean_hos, std_hos=std_hos, nvalues=nvalues, oscan=oscan.copy())
        chans.append(d)

    for i in range(nchans):
        header['BIASM{}'.format(i + 1)] = mean_vos[i]
    for i in range(nchans):
        header['RDN{}'.format(i + 1)] = std_vos[i]
    header['BIASMEAN'] = np.nanmean(mean_vos)
    header['RDNOISE'] = np.nanmean(std_vos)
    if diag is not None:
        diag['chans'] = chans
        diag['mean_vos'] = mean_vos
        diag['std_vos'] = std_vos
    return data_out


# -------------------------------------------------------------------------------------------
def fill_sat_holes(data_mask, mask_value):
    """blackbox.py:4584-4596"""
    vs, vc = mask_value['saturated'], mask_value['saturated-connected']
    m = (data_mask & vs == vs) | (data_mask & vc == vc)
    struct = np.ones((3, 3), dtype=bool)
    m = ndimage.binary_closing(m, structure=struct)
    m = ndimage.binary_fill_holes(m, structure=struct)
    data_mask[m & (data_mask == 0)] = vc


def mask_init(data, header, bpm, imgtype, tel=None, diag=None):
    """Initial mask; blackbox.py:4375-4579 with the bad-pixel-mask FITS file replaced by the
    array ``bpm`` (None = no BPM).  Returns (uint8 mask, mask header dict); ``data`` has its
    non-finite pixels zeroed in place."""
    data_mask = (np.zeros(data.shape, dtype='uint8') if bpm is None
                 else np.array(bpm, dtype='uint8', copy=True))
    header_mask = {}
    mask_value = get_par(set_bb.mask_value, tel)
    if imgtype == 'object':
        infnan = ~np.isfinite(data)
        data[infnan] = 0
        data_mask[infnan & (data_mask == 0)] |= mask_value['bad']
        data_sec_red = define_sections(data.shape, tel=tel)[4]
        nchans = len(data_sec_red)
        biaslevel = np.array([header['BIASM{}'.format(i + 1)] for i in range(nchans)])
        satlevel_chans = (np.array(get_par(set_bb.satlevel, tel)) *
                          np.array(get_par(set_bb.gain, tel)) - biaslevel)
        header_mask['SATURATE'] = header['SATURATE'] = np.mean(satlevel_chans)
        mask_sat = np.zeros(data.shape, dtype=bool)
        for i in range(nchans):
            key = 'SATLEV{}'.format(i + 1)
            header[key] = header_mask[key] = round(satlevel_chans[i], 1)
            sat_i = data[data_sec_red[i]] >= satlevel_chans[i]
            mask_sat[data_sec_red[i]] = sat_i
            sat_i_flip = np.flipud(sat_i)
            for v in range(nchans):
                if v != i:
                    use = sat_i if i // 8 == v // 8 else sat_i_flip
                    data_mask[data_sec_red[v]][use] |= mask_value['crosstalk']
        data_mask[mask_sat] |= mask_value['saturated']
        struct = np.ones((3, 3), dtype=bool)
        nobj = ndimage.label(mask_sat, structure=struct)[1]
        header_mask['NOBJ-SAT'] = header['NOBJ-SAT'] = nobj
        satcon = ndimage.binary_dilation(mask_sat, structure=struct, iterations=1)
        data_mask[satcon & ~mask_sat] |= mask_value['saturated-connected']
        if diag is not None:
            diag['mask_before_fill'] = data_mask.copy()
            diag['mask_sat'] = mask_sat
        fill_sat_holes(data_mask, mask_value)
    return data_mask.astype('uint8'), header_mask


def mask_header(data_mask, header_mask, tel=None):
    """blackbox.py:4601-4620: M-<type>, M-<type>VAL, M-<type>NUM per mask type."""
    text = {'bad': 'BP', 'edge': 'EP', 'saturated': 'SP', 'saturated-connected': 'SCP',
            'satellite trail': 'STP', 'cosmic ray': 'CRP'}
    mask_value = get_par(set_bb.mask_value, tel)
    for mask_type, short in text.items():
        value = mask_value[mask_type]
        header_mask['M-' + short] = True
        header_mask['M-{}VAL'.format(short)] = value
        header_mask['M-{}NUM'.format(short)] = int(np.sum(data_mask & value == value))


# -------------------------------------------------------------------------------------------
def cosmics_corr(data, header, data_mask, header_mask, tel=None, niter=None):
    """blackbox.py:4259-4370; returns (cleaned data, mask with the cosmic-ray bit)."""
    mask_cr, data = lacosmic.detect_cosmics(
        data, inmask=(data_mask != 0),
        sigclip=get_par(set_bb.sigclip, tel), sigfrac=get_par(set_bb.sigfrac, tel),
        objlim=get_par(set_bb.objlim, tel),
        niter=get_par(set_bb.niter, tel) if niter is None else niter,
        readnoise=header['RDNOISE'], gain=1.0, satlevel=np.inf, cleantype='medmask',
        sepmed=get_par(set_bb.sepmed, tel))
    data_mask[mask_cr == 1] |= get_par(set_bb.mask_value, tel)['cosmic ray']
    ncosmics = ndimage.label(mask_cr, structure=np.ones((3, 3), dtype=bool))[1]
    header['NCOSMICS'] = header_mask['NCOSMICS'] = ncosmics / float(header['EXPTIME'])
    return data, data_mask


# -------------------------------------------------------------------------------------------
def xtalk_coeffs(victim, source, correction, nchans=16):
    """coefficient matrix [source, victim] from the 1-based table columns;
    blackbox.py:7159-7198"""
    coeffs = np.zeros((nchans, nchans))
    for v, s, c in zip(victim, source, correction):
        coeffs[int(s) - 1, int(v) - 1] = c
    return coeffs


def xtalk_corr(data, coeffs, data_mask=None, tel=None):
    """In-place crosstalk correction; blackbox.py:7138-7258 with the ASCII table already
    parsed into ``coeffs[source, victim]``."""
    if data_mask is None:
        data_mask = np.zeros(data.shape, dtype=bool)
    chan_sec = define_sections(data.shape, tel=tel)[0]
    nchans = len(chan_sec)
    mv = get_par(set_bb.mask_value, tel)
    mask_source = ((data > 0) & (data_mask & mv['bad'] == 0) &
                   (data_mask & mv['cosmic ray'] == 0))
    mask_victim = (data_mask & mv['edge'] == 0)
    stack = np.stack([data[s] * mask_source[s] for s in chan_sec], axis=2)
    stack_flip = np.stack([np.flipud(data[s] * mask_source[s]) for s in chan_sec], axis=2)
    ysize, xsize = data[chan_sec[0]].shape
    corr = np.zeros((nchans, ysize, xsize))
    s1, s2 = slice(0, 8), slice(8, 16)
    for q, (sls, slv) in enumerate([(s1, s1), (s2, s1), (s1, s2), (s2, s2)]):
        use = stack if q in (0, 3) else stack_flip
        corr[slv] += np.matmul(use[:, :, sls], coeffs[sls, slv]).swapaxes(0, 2).swapaxes(1, 2)
    for i in range(nchans):
        data[chan_sec[i]] -= corr[i] * mask_victim[chan_sec[i]]


# -------------------------------------------------------------------------------------------
def master_median(frames, imgtype='bias', medsec=None, bpm=None, tel=None):
    """Arithmetic core of master_prep (blackbox.py:4908-4984, 5063-5073): stack, for flats
    divide frame i by its normalisation median (``medsec[i]`` = header MEDSEC, or the median
    over set_bb.flat_norm_sec), np.median along the stack, flats: edge | <=0 -> 1."""
    nfiles = len(frames)
    cube = np.zeros((nfiles,) + frames[0].shape, dtype=F32)
    scales = []
    for i, f in enumerate(frames):
        cube[i] = f
        if imgtype == 'flat':
            if medsec is not None and medsec[i] is not None:
                median = medsec[i]
            else:
                median = np.median(cube[i][get_par(set_bb.flat_norm_sec, tel)])
            scales.append(median)
            if median != 0:
                cube[i] /= median
    out = np.median(cube, axis=0)
    if imgtype == 'flat' and bpm is not None:
        out[(bpm == get_par(set_bb.mask_value, tel)['edge']) | (out <= 0)] = 1
    return out, scales


def master_flat_stats(frames, medsec=None, bpm=None, tel=None):
    """The deterministic part of the master-flat header (blackbox.py:5006-5013, 5081-5161):
    MFMEDSEC / MFSTDSEC over flat_norm_sec of the median BEFORE the edge / non-positive fix, and
    the channel factors GAINCF of the fixed master.  -> (master, dict)."""
    master_median, _ = master_median_unfixed(frames, medsec, tel)
    sec_tmp = get_par(set_bb.flat_norm_sec, tel)
    out = {'MFMEDSEC': np.median(master_median[sec_tmp]), 'MFSTDSEC': np.std(master_median[sec_tmp])}
    if bpm is not None:
        master_median[(bpm == get_par(set_bb.mask_value, tel)['edge']) | (master_median <= 0)] = 1
    data_shape = master_median.shape
    data_sec_red = define_sections(data_shape, tel=tel)[4]
    nchans = np.shape(data_sec_red)[0]
    med_chan_cntr = np.zeros(nchans)
    master_median_corr = np.copy(master_median)
    nrows = 200
    for i_chan in range(nchans):
        data_chan = master_median_corr[data_sec_red[i_chan]]
        if i_chan < 8:
            med_chan_cntr[i_chan] = np.median(data_chan[-nrows:, :])
        else:
            med_chan_cntr[i_chan] = np.median(data_chan[0:nrows, :])
        master_median_corr[data_sec_red[i_chan]] /= med_chan_cntr[i_chan]
    factor_chan = 1. / med_chan_cntr
    ysize, xsize = data_shape
    ny, nx = get_par(set_bb.ny, tel), get_par(set_bb.nx, tel)
    dy, dx = ysize // ny, xsize // nx
    nrows, ncols = 2000, 200
    for i in range(1, nx):
        y_index, x_index = dy, i * dx
        data_stat1 = master_median_corr[y_index - nrows:y_index + nrows, x_index - ncols:x_index]
        data_stat2 = master_median_corr[y_index - nrows:y_index + nrows, x_index:x_index + ncols]
        ratio = np.median(data_stat1) / np.nanmedian(data_stat2)
        master_median_corr[data_sec_red[i]] *= ratio
        master_median_corr[data_sec_red[i + nx]] *= ratio
        factor_chan[i] *= ratio
        factor_chan[i + nx] *= ratio
    factor_chan /= np.mean(factor_chan)
    for i_chan in range(nchans):
        out['GAINCF{}'.format(i_chan + 1)] = factor_chan[i_chan]
    return master_median, out


def master_median_unfixed(frames, medsec=None, tel=None):
    return master_median(frames, imgtype='flat', medsec=medsec, bpm=None, tel=tel)


def nonlin_corr(data, fit_splines, tel=None):
    """In place; blackbox.py:7392-7437 verbatim (fit_splines: the unpickled list of 16 spline
    objects).  Note the reference's ``frac_corr = np.ones(...)``: above 50000 counts the data are
    divided by 2."""
    gain = get_par(set_bb.gain, tel)
    data_sec_red = define_sections(np.shape(data), tel=tel)[4]
    for i_chan in range(len(data_sec_red)):
        data_counts = data[data_sec_red[i_chan]] / gain[i_chan]
        frac_corr = np.ones(data_counts.shape)
        with np.errstate(invalid='ignore'):
            mask_corr = (data_counts <= 50000)
        frac_corr[mask_corr] = fit_splines[i_chan](data_counts[mask_corr])
        data[data_sec_red[i_chan]] /= (frac_corr + 1)
    return data


def master_median_clipped(frames, sigma=3.0, maxiters=5, scales=None):
    """Sigma-clipped median combine (NOT the reference's master_prep, which takes the plain
    median, blackbox.py:4984; BASELINE.json's wording): astropy.stats.sigma_clip along the stack
    axis with cenfunc='median', then np.ma.median; NaN where nothing survives."""
    cube = np.stack([np.asarray(f, dtype=F32) for f in frames])
    if scales is not None:
        for i, sc in enumerate(scales):
            if sc != 0:
                cube[i] /= F32(sc)
    clipped = sigma_clip(cube, sigma=sigma, maxiters=maxiters, cenfunc='median', axis=0, masked=True)
    med = np.ma.median(clipped, axis=0)
    return np.ma.filled(med.astype(F32), np.nan).astype(F32)


def fill_edge_pixels(data, data_mask, tel=None):
    """In place: edge pixels -> median of their channel; blackbox.py:1958-1974."""
    value_edge = get_par(set_bb.mask_value, tel)['edge']
    mask_edge = (data_mask & value_edge == value_edge)
    data_sec_red = define_sections(data.shape, tel=tel)[4]
    meds = []
    for sec in data_sec_red:
        med = np.median(data[sec])
        meds.append(med)
        data[sec][mask_edge[sec]] = med
    return np.array(meds, dtype=F32)


# -------------------------------------------------------------------------------------------
def reduce_frame(raw, tel, mbias=None, mflat=None, bpm=None, coeffs=None, exptime=60.0,
                 niter=None, steps=('gain', 'os', 'bias', 'mask', 'flat', 'cosmics', 'xtalk'),
                 diag=None):
    """One science frame through the chain in blackbox_reduce's order
    (blackbox.py:1451-1902).  raw: uint16 or float32 raw frame with overscans."""
    header = {'EXPTIME': exptime}
    data = np.array(raw, dtype=F32)             # read_hdulist(dtype='float32')
    data[~np.isfinite(data)] = 0
    if 'gain' in steps:
        gain_corr(data, header, tel=tel)
    data = os_corr(data, header, 'object', tel=tel, diag=diag)
    if 'bias' in steps and mbias is not None and get_par(set_bb.subtract_mbias, tel):
        data -= mbias
    data_mask, header_mask = None, {}
    if 'mask' in steps:
        data_mask, header_mask = mask_init(data, header, bpm, 'object', tel=tel)
    if 'flat' in steps and mflat is not None:
        data /= mflat
    if 'cosmics' in steps:
        data, data_mask = cosmics_corr(data, header, data_mask, header_mask, tel=tel,
                                       niter=niter)
    if 'xtalk' in steps and coeffs is not None:
        xtalk_corr(data, coeffs, data_mask, tel=tel)
    return data, data_mask, header, header_mask
