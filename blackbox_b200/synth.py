"""Seeded synthetic MeerLICHT/BlackGEM frames.

Everything is ``numpy.random.default_rng(seed)`` on the host so the same bits feed the CPU
oracle and the GPU path.  Shapes are parametric (``ysize_chan`` / ``xsize_chan`` / overscan
sizes) so parity tests can run on frames of a few Mpx while the benchmark uses the full
10600 x 12000 raw frame (data area 10560 x 10560, 16 channels of 5280 x 1320, 20 horizontal-
overscan rows and 180 vertical-overscan columns per channel).

Raw layout per channel tile (dy x dx): bottom half -- data rows [0, ysize_chan), then the
horizontal overscan; top half mirrored in y (overscan rows first).  Columns: data
[0, xsize_chan), then the vertical overscan.
"""
import numpy as np

from . import set_bb
from .set_bb import get_par

BIAS_ADU = {'ML1': 3050.0, 'BG': 1200.0}


def raw_shape(ysize_chan=None, xsize_chan=None, os_rows=20, os_cols=180):
    ysc = set_bb.ysize_chan if ysize_chan is None else ysize_chan
    xsc = set_bb.xsize_chan if xsize_chan is None else xsize_chan
    return (set_bb.ny * (ysc + os_rows), set_bb.nx * (xsc + os_cols))


def _add_stars(img, rng, nstars, fwhm_range=(2.5, 4.0), flux_min=300.0, slope=1.6,
               flux_max=3.0e7):
    """Gaussian stars with a power-law flux distribution, added in place (units of img)."""
    H, W = img.shape
    u = rng.random(nstars)
    a = 1.0 - slope
    flux = (flux_min ** a + u * (flux_max ** a - flux_min ** a)) ** (1.0 / a)
    ys = rng.uniform(0, H, nstars)
    xs = rng.uniform(0, W, nstars)
    fwhm = rng.uniform(*fwhm_range, nstars)
    for f, y0, x0, fw in zip(flux, ys, xs, fwhm):
        sig = fw / 2.3548
        r = int(min(max(4 * sig + 2, np.sqrt(max(f, 1.0)) * 0.02 + 4 * sig), 60))
        ya, yb = max(int(y0) - r, 0), min(int(y0) + r + 1, H)
        xa, xb = max(int(x0) - r, 0), min(int(x0) + r + 1, W)
        if ya >= yb or xa >= xb:
            continue
        yy = np.arange(ya, yb, dtype=np.float32)[:, None] - np.float32(y0)
        xx = np.arange(xa, xb, dtype=np.float32)[None, :] - np.float32(x0)
        img[ya:yb, xa:xb] += (f / (2 * np.pi * sig * sig)) * np.exp(
            -(yy * yy + xx * xx) / (2 * sig * sig))
    return flux


def _add_cosmics(img, rng, ncosmics, amp_range=(200.0, 50000.0), max_len=12):
    """Sharp cosmic-ray tracks (1..max_len px), added in place; returns the hit mask."""
    H, W = img.shape
    hit = np.zeros(img.shape, dtype=bool)
    for _ in range(ncosmics):
        y, x = rng.uniform(3, H - 3), rng.uniform(3, W - 3)
        length = int(rng.integers(1, max_len + 1))
        ang = rng.uniform(0, np.pi)
        amp = np.exp(rng.uniform(np.log(amp_range[0]), np.log(amp_range[1])))
        for t in range(length):
            yi, xi = int(round(y + t * np.sin(ang))), int(round(x + t * np.cos(ang)))
            if 0 <= yi < H and 0 <= xi < W:
                img[yi, xi] += amp * rng.uniform(0.5, 1.0)
                hit[yi, xi] = True
    return hit


def make_sky(tel, seed, ysize_chan=None, xsize_chan=None, nstars=None, ncosmics=None,
             sky_adu=150.0, level_adu=None):
    """Photon image of the data area in ADU (float32, no bias, no flat), with stars and
    cosmic-ray hits; also returns the cosmic-ray truth mask.  ``level_adu`` makes a flat
    field exposure of that level instead of a star field."""
    rng = np.random.default_rng(seed)
    ysc = set_bb.ysize_chan if ysize_chan is None else ysize_chan
    xsc = set_bb.xsize_chan if xsize_chan is None else xsize_chan
    H, W = set_bb.ny * ysc, set_bb.nx * xsc
    area_frac = (H * W) / (10560.0 * 10560.0)
    if level_adu is not None:
        img = np.full((H, W), np.float32(level_adu), dtype=np.float32)
        nstars, ncosmics = 0, (0 if ncosmics is None else ncosmics)
    else:
        img = np.full((H, W), np.float32(sky_adu), dtype=np.float32)
        if nstars is None:
            nstars = max(int(5000 * area_frac), 20)
        if ncosmics is None:
            ncosmics = max(int(2000 * area_frac), 10)
    if nstars:
        _add_stars(img, rng, nstars)
    # photon noise (normal approximation, fine for >= 100 ADU)
    noise = rng.standard_normal(img.shape, dtype=np.float32)
    img += np.sqrt(np.maximum(img, 1.0) / 2.2, dtype=np.float32) * noise
    hit = _add_cosmics(img, rng, ncosmics) if ncosmics else np.zeros(img.shape, bool)
    return img, hit


def make_flat_response(seed, shape, pix_noise=0.01):
    """Vignetting 1 - 0.1 r^2 times 1 % pixel-to-pixel response (float32, ~1)."""
    rng = np.random.default_rng(seed)
    H, W = shape
    yy = (np.arange(H, dtype=np.float32)[:, None] - H / 2) / (H / 2)
    xx = (np.arange(W, dtype=np.float32)[None, :] - W / 2) / (W / 2)
    resp = (1.0 - 0.05 * (yy * yy + xx * xx)).astype(np.float32)
    resp *= (1.0 + pix_noise * rng.standard_normal(shape, dtype=np.float32))
    return resp.astype(np.float32)


def make_raw(tel, seed, ysize_chan=None, xsize_chan=None, os_rows=20, os_cols=180,
             sky=None, response=None, read_noise_adu=4.0, **sky_kw):
    """Raw uint16 frame with overscans.  Returns (raw, truth dict)."""
    rng = np.random.default_rng(seed + 7919)
    ysc = set_bb.ysize_chan if ysize_chan is None else ysize_chan
    xsc = set_bb.xsize_chan if xsize_chan is None else xsize_chan
    ny, nx = set_bb.ny, set_bb.nx
    dy, dx = ysc + os_rows, xsc + os_cols
    cosmic_truth = None
    if sky is None:
        sky, cosmic_truth = make_sky(tel, seed, ysc, xsc, **sky_kw)
    if response is not None:
        sky = sky * response
    base = get_par(BIAS_ADU, tel)
    bias = base + 30.0 * rng.standard_normal(ny * nx)
    raw = np.empty((ny * dy, nx * dx), dtype=np.float32)
    yrow = np.linspace(-1.0, 1.0, dy, dtype=np.float32)
    xcol = np.arange(dx, dtype=np.float32)
    for i in range(ny * nx):
        r, c = divmod(i, nx)
        drift = 2.0 * rng.uniform(-1, 1, 4)
        vos = (drift[0] * yrow ** 3 + drift[1] * yrow ** 2 + drift[2] * yrow).astype(np.float32)
        hos = (5.0 * np.exp(-xcol / 40.0) + rng.uniform(-1, 1) * 1e-9 * (xcol - 700.0) ** 3
               ).astype(np.float32)
        tile = np.float32(bias[i]) + vos[:, None] + hos[None, :]
        tile = tile + np.float32(read_noise_adu) * rng.standard_normal((dy, dx), dtype=np.float32)
        photons = sky[r * ysc:(r + 1) * ysc, c * xsc:(c + 1) * xsc]
        if r == 0:
            tile[:ysc, :xsc] += photons
        else:
            tile[os_rows:, :xsc] += photons
        raw[r * dy:(r + 1) * dy, c * dx:(c + 1) * dx] = tile
    raw = np.clip(np.rint(raw), 0, 65535).astype(np.uint16)
    return raw, {'bias_adu': bias, 'cosmics': cosmic_truth}


def make_masters(tel, seed, shape, edge=20, bad_frac=1e-3):
    """(master bias f32 ~N(0,1) e-, master flat f32 ~1, bad-pixel mask u8)."""
    rng = np.random.default_rng(seed + 104729)
    mbias = rng.standard_normal(shape, dtype=np.float32)
    mflat = make_flat_response(seed + 1, shape)
    mflat /= np.float32(np.median(mflat))
    mv = set_bb.mask_value
    bpm = np.zeros(shape, dtype=np.uint8)
    bpm[rng.random(shape, dtype=np.float32) < bad_frac] = mv['bad']
    e = min(edge, shape[0] // 8, shape[1] // 8)
    if e > 0:
        bpm[:e, :] = mv['edge']
        bpm[-e:, :] = mv['edge']
        bpm[:, :e] = mv['edge']
        bpm[:, -e:] = mv['edge']
    xsc = shape[1] // set_bb.nx
    for c in range(1, set_bb.nx):                    # channel-edge columns
        bpm[:, c * xsc - 1:c * xsc + 1] |= mv['bad']
    return mbias, mflat, bpm


def make_xtalk(seed, nchans=16, amp=3e-4):
    """All 240 off-diagonal (victim, source, correction) rows, 1-based channel numbers, and
    the coefficient matrix [source, victim] they define (blackbox.py:7159-7198)."""
    rng = np.random.default_rng(seed + 15485863)
    victim, source, corr = [], [], []
    for v in range(1, nchans + 1):
        for s in range(1, nchans + 1):
            if v != s:
                victim.append(v)
                source.append(s)
                corr.append(rng.uniform(-amp, amp))
    coeffs = np.zeros((nchans, nchans))
    for v, s, c in zip(victim, source, corr):
        coeffs[s - 1, v - 1] = c
    return np.array(victim), np.array(source), np.array(corr), coeffs


def write_xtalk_file(path, victim, source, corr):
    """ASCII table with a header line, as the reference's new crosstalk files
    (blackbox.py:7156-7161)."""
    with open(path, 'w') as fh:
        fh.write('victim source correction\n')
        for v, s, c in zip(victim, source, corr):
            fh.write('{} {} {!r}\n'.format(int(v), int(s), float(c)))


# -------------------------------------------------------------------------------------------
# a night's worth of reduced calibration frames for master_prep (blackbox.py:4625-5247)
# -------------------------------------------------------------------------------------------
def make_cal_night(tel, imgtype, seed, shape, date_eve='20240105', filt='q', with_data=True):
    """-> list of (relative path below red_dir, float32 frame, header dict) of reduced bias /
    flat frames around the evening date ``date_eve``, with what master_prep's selection has to
    deal with: a red-flagged frame, evening flats, frames of neighbouring nights, flats with and
    without MEDSEC, a cluster of non-positive pixels inside the statistics section.
    ``with_data=False``: names and headers only (frames None, no MEDSEC), for the selection logic."""
    import datetime
    rng = np.random.default_rng(seed)
    d0 = datetime.datetime.strptime(date_eve, '%Y%m%d')
    mjd_eve = (d0 - datetime.datetime(1858, 11, 17)).days          # MJD at 00:00 of the evening date
    out = []
    if imgtype == 'flat':
        chan_level = 1.0 + 0.02 * rng.standard_normal(16)
        resp = make_flat_response(seed + 1, shape, pix_noise=0.005) if with_data else None
        ysc, xsc = shape[0] // 2, shape[1] // 8
        for c in range(16 if with_data else 0):
            resp[(c // 8) * ysc:(c // 8 + 1) * ysc, (c % 8) * xsc:(c % 8 + 1) * xsc] *= np.float32(chan_level[c])
        sec = set_bb.get_par(set_bb.flat_norm_sec, tel)
        # morning flats of the next UT day (fraction 0.3-0.45), two evening flats, one red-flagged
        plan = [(1, 0.30 + 0.02 * k, None) for k in range(7)] + [(0, 0.75, None), (1, 0.05, None), (1, 0.44, 'red')]
        for k, (day, frac, flag) in enumerate(plan):
            level = 20000.0 * (1.0 + 0.2 * rng.uniform(-1, 1))
            frame = None
            if with_data:
                frame = (np.float32(level) * resp * (1.0 + 0.004 * rng.standard_normal(shape, dtype=np.float32))
                         ).astype(np.float32)
                frame[sec[0].start + 10:sec[0].start + 13, sec[1].start + 20:sec[1].start + 24] = -5.0
                frame[5, 7] = 0.0
            mjd = mjd_eve + day + frac
            t = datetime.datetime(1858, 11, 17) + datetime.timedelta(days=mjd)
            hdr = {'IMAGETYP': 'flat', 'FILTER': filt, 'MJD-OBS': float(mjd), 'DATE-OBS': t.isoformat(),
                   'RA': 150.0 + 0.004 * k * (k % 3 != 0), 'DEC': -30.0 + 0.003 * k, 'ORIGFILE': 'raw_{:03d}'.format(k)}
            if k % 2 == 0 and with_data:
                hdr['MEDSEC'] = float(np.median(frame[sec])) * 1.0005      # the header value wins
            if flag:
                hdr['QC-FLAG'] = flag
            sub = d0.strftime('%Y/%m/%d')
            name = '{}/flat/{}_{}_{}_red_{}.fits'.format(sub, tel, t.strftime('%Y%m%d'), t.strftime('%H%M%S'), filt)
            out.append((name, frame, hdr))
        return out
    # biases: 23 frames over three nights, one red-flagged; ncal_max = 20 keeps the nearest
    for k in range(23):
        night = (-1, 0, 1)[k % 3]
        frac = 0.60 + 0.01 * k if k % 2 else 0.20 + 0.01 * k
        mjd = mjd_eve + night + frac
        t = datetime.datetime(1858, 11, 17) + datetime.timedelta(days=mjd)
        frame = (3.0 * rng.standard_normal(shape, dtype=np.float32) + np.float32(0.1 * k)).astype(np.float32) \
            if with_data else None
        hdr = {'IMAGETYP': 'bias', 'MJD-OBS': float(mjd), 'DATE-OBS': t.isoformat(), 'ORIGFILE': 'raw_{:03d}'.format(k)}
        if k == 4:
            hdr['QC-FLAG'] = 'red'
        sub = (d0 + datetime.timedelta(days=night)).strftime('%Y/%m/%d')
        name = '{}/bias/{}_{}_{}_red.fits'.format(sub, tel, t.strftime('%Y%m%d'), t.strftime('%H%M%S'))
        out.append((name, frame, hdr))
    return out


def add_hos_contamination(raw, ysize_chan, os_rows=20, xsize_chan=None, os_cols=180, level=3000):
    """Bright charge leaking into the horizontal overscan (what MeerLICHT's ``data_limit`` mask in
    os_corr is for, blackbox.py:6586-6614): in channels 3 (bottom) and 12 (top) one isolated
    column and a six-column band get ``level`` ADU more over the overscan rows and the data rows
    next to them, and one isolated column in only 3 of the overscan rows (less than half the
    height: it stays masked pixel by pixel).  In place."""
    xs = set_bb.xsize_chan if xsize_chan is None else xsize_chan
    wchan = xs + os_cols
    hchan = ysize_chan + os_rows
    for chan, rows in ((2, slice(hchan - os_rows - 30, hchan)), (11, slice(hchan, hchan + os_rows + 30))):
        x0 = (chan % 8) * wchan
        for cols in (slice(x0 + 400, x0 + 401), slice(x0 + 700, x0 + 706)):
            raw[rows, cols] = np.minimum(raw[rows, cols].astype(np.int64) + level, 65535).astype(raw.dtype)
        short = slice(hchan - 8, hchan - 5) if chan < 8 else slice(hchan + 5, hchan + 8)
        raw[short, x0 + 900] = np.minimum(raw[short, x0 + 900].astype(np.int64) + level, 65535).astype(raw.dtype)
    return raw


def add_saturated_rings(raw, centres=((60, 700), (120, 4000), (150, 9100), (300, 2500))):
    """Saturated rings with an unsaturated inside (and one broken ring, and one ring with a
    saturated island in it): what fill_sat_holes has to close and fill (blackbox.py:4584-4596).
    In place; positions are raw-frame pixels inside data sections of a frame with ysize_chan >= 200."""
    yy, xx = np.mgrid[-16:17, -16:17]
    rr = np.hypot(yy, xx)
    for n, (cy, cx) in enumerate(centres):
        ring = (rr >= 9.5) & (rr <= 12.5)
        if n == 1:
            ring &= ~((xx > 0) & (np.abs(yy) < 3))              # a gap wider than the closing: stays open
        if n == 2:
            ring |= rr <= 1.5                                    # an island inside the hole
        if n == 3:
            ring &= ~((xx > 0) & (np.abs(yy) < 1))              # a one-pixel gap: closed by the 3x3 closing
        sub = raw[cy - 16:cy + 17, cx - 16:cx + 17]
        sub[ring] = 65535
    return raw


def add_nonfinite(data, bpm):
    """NaN / inf pixels in an overscan-corrected frame, two of them on pixels the bad-pixel mask
    already flags: mask_init zeroes them and marks the unflagged ones 'bad'
    (blackbox.py:4405-4413).  In place."""
    data[100, 100:104] = np.nan
    data[50, 5000] = np.inf
    data[51, 5000] = -np.inf
    ys, xs = np.nonzero(bpm[60:140, 2000:3000])
    for y, x in list(zip(ys, xs))[:2]:
        data[60 + y, 2000 + x] = np.nan
    return data
