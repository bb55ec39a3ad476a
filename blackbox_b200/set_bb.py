"""Hot-path configuration constants.

Mirror of the attribute names the reference reads through ``get_par(set_bb.X, tel)``
(reference: Settings/set_blackbox.py:36-49 [subtract_mbias, ncal_max, voscan_poldeg],
:211-218 [LACosmic], :241-311 [gain, satlevel], :324-337 [flat_norm_sec, ny, nx,
ysize_chan, xsize_chan]).  Only the names the reduction hot path touches exist here.

``mask_value`` lives in the reference's *set_zogy* module, which is not in the reference
tree; the values are ZOGY's published defaults and stay configurable.
"""

# subtract master bias (reference: set_blackbox.py:37)
subtract_mbias = {'ML1': False, 'BG': True}

# maximum number of calibration frames combined into a master (set_blackbox.py:49)
ncal_max = {'bias': 20, 'dark': 20, 'flat': 15}

# days either side of the evening date whose calibration frames go into a master
# (set_blackbox.py:47)
cal_window = {'bias': 3, 'dark': 3, 'flat': 7}

# BlackGEM evening flats show a reflection and are not used (set_blackbox.py:331)
flat_reject_eve = {'ML': False, 'BG': True}

# site folders read by master_prep (set_blackbox.py:89-152 derive them from the processing
# environment; here plain per-telescope settings, to be pointed at the site's folders):
# reduced frames live in [red_dir]/yyyy/mm/dd/[imgtype]/, masters in
# [master_dir]/yyyy/mm/dd/[imgtype]/ (blackbox.py:1127-1128, 1649, 1797)
red_dir = {'ML1': '/data/red/ML1', 'BG2': '/data/red/BG2', 'BG3': '/data/red/BG3', 'BG4': '/data/red/BG4'}
master_dir = {'ML1': '/data/masters/ML1', 'BG2': '/data/masters/BG2', 'BG3': '/data/masters/BG3',
              'BG4': '/data/masters/BG4'}

# bad-pixel masks; master_prep and mask_init insert the filter: 'bpm' -> 'bpm_<filt>'
# (set_blackbox.py:187-193)
cal_dir = '/data/CalFiles'
bad_pixel_mask = {'ML1': cal_dir + '/BPM/ML1/ML1_bpm_0p2_20200727.fits',
                  'BG2': cal_dir + '/BPM/BG2/BG2_bpm_0p2_20250130.fits',
                  'BG3': cal_dir + '/BPM/BG3/BG3_bpm_0p2_20250130.fits',
                  'BG4': cal_dir + '/BPM/BG4/BG4_bpm_0p2_20250130.fits'}

# degree of the polynomial fitted to the vertical-overscan row means (set_blackbox.py:52)
voscan_poldeg = 3

# LACosmic parameters handed to detect_cosmics (set_blackbox.py:211-218)
sigclip = {'ML1': 15, 'BG': 20}
sigfrac = 0.01
objlim = 3
niter = 3
sepmed = False

# channel gains [e-/ADU]; index = 8*row + col, row 0 = bottom (set_blackbox.py:241-281)
gain = {
    'ML1': [2.112, 2.125, 2.130, 2.137, 2.156, 2.158, 2.163, 2.164,
            2.109, 2.124, 2.126, 2.132, 2.136, 2.154, 2.155, 2.157],
    'BG2': [2.694, 2.685, 2.691, 2.661, 2.655, 2.673, 2.695, 2.659,
            2.654, 2.748, 2.712, 2.717, 2.714, 2.702, 2.673, 2.743],
    'BG3': [2.614, 2.609, 2.634, 2.647, 2.600, 2.616, 2.683, 2.649,
            2.680, 2.679, 2.644, 2.604, 2.615, 2.633, 2.615, 2.714],
    'BG4': [2.415, 2.393, 2.365, 2.333, 2.340, 2.320, 2.348, 2.389,
            2.395, 2.403, 2.381, 2.350, 2.362, 2.369, 2.391, 2.430],
}

# channel saturation levels [ADU] of raw images (set_blackbox.py:296-305)
satlevel = {
    'ML1': [5.89e4, 5.94e4, 5.82e4, 5.59e4, 5.60e4, 5.63e4, 5.60e4, 5.75e4,
            5.88e4, 5.81e4, 5.71e4, 5.65e4, 5.59e4, 5.60e4, 5.59e4, 5.65e4],
    'BG2': [3.84e4, 3.77e4, 3.75e4, 3.79e4, 3.79e4, 3.80e4, 3.75e4, 3.93e4,
            4.50e4, 4.08e4, 4.08e4, 4.09e4, 4.07e4, 3.95e4, 4.15e4, 4.37e4],
    'BG3': [3.96e4, 3.83e4, 3.79e4, 3.77e4, 3.81e4, 3.83e4, 3.74e4, 3.94e4,
            4.00e4, 3.98e4, 4.13e4, 4.29e4, 4.29e4, 4.22e4, 4.13e4, 4.38e4],
    'BG4': [4.11e4, 4.09e4, 4.16e4, 4.29e4, 4.32e4, 4.29e4, 4.23e4, 4.41e4,
            4.66e4, 4.60e4, 4.53e4, 4.67e4, 4.66e4, 4.65e4, 4.64e4, 4.66e4],
}

# reduced-image section whose median normalises a flat (set_blackbox.py:324-327)
flat_norm_sec = {'ML1': (slice(6600, 9240), slice(5280, 7920)),
                 'BG2': (slice(500, 2000), slice(1320, 6600)),
                 'BG3': (slice(300, 1200), slice(5280, 10000)),
                 'BG4': (slice(2640, 5280), slice(3960, 7920))}

# number of channels in y and x, and the data-section size per channel (set_blackbox.py:335-337)
ny, nx = 2, 8
ysize_chan, xsize_chan = 5280, 1320

# BlackGEM row windows next to the horizontal overscan in which saturated pixels flag a
# column as leaking into the overscan (reference: blackbox.py:6625)
hos_sat_ypix_lim = {'BG2': (2640, 5280), 'BG3': (1320, 2640), 'BG4': (1320, 2640)}

# set_zogy.mask_value (ZOGY defaults; the module itself is absent from the reference tree)
mask_value = {'bad': 1, 'cosmic ray': 2, 'saturated': 4, 'saturated-connected': 8,
              'satellite trail': 16, 'edge': 32, 'crosstalk': 64}


def get_par(par, tel):
    """Resolve a telescope-keyed setting.

    Same rule as the reference helper (copy at buildref.py:3889-3906): exact key first,
    then the alphabetic prefix of ``tel`` (``'BG3'`` -> ``'BG'``); non-dict values pass
    through untouched.
    """
    if isinstance(par, dict):
        if tel in par:
            return par[tel]
        base = ''.join(c for c in str(tel) if c.isalpha())
        if base in par:
            return par[base]
    return par
