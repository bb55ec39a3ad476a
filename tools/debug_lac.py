#!/usr/bin/env python
"""Development aid: where do the lazy / dense LACosmic paths or two runs of the chain differ?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from blackbox_b200 import reduce as bbr, set_bb, synth  # noqa: E402
from blackbox_b200.pipeline import FramePipeline  # noqa: E402


def stress():
    import test_steps_gpu as T
    from oracle import lacosmic
    img, mask = T._lacosmic_case(11, shape=(1500, 2000), ncr=6000, masked_frac=0.02)
    for niter in (3, 4, 5):
        kw = dict(sigclip=5.0, sigfrac=0.01, objlim=2, niter=niter, readnoise=8.5, gain=1.0,
                  satlevel=np.inf, cleantype='medmask', sepmed=False)
        io = {}
        cr_o, clean_o = lacosmic.detect_cosmics(img, inmask=mask, info=io, **kw)
        for mode, name in ((bbr.LAC_LAZY, 'lazy'), (bbr.LAC_LAZY, 'lazy2'), (bbr.LAC_LAZY_BG, 'lazybg'), (bbr.LAC_DENSE, 'dense')):
            ig = {}
            cr_g, clean_g = bbr.detect_cosmics(img, inmask=mask, info=ig, mode=mode, **kw)
            d = np.argwhere(cr_g != cr_o)
            dc = np.argwhere(clean_g.view(np.uint32) != clean_o.view(np.uint32))
            print('niter', niter, name, 'ncr', list(ig['ncr_per_iter']), 'oracle', list(io['ncr_per_iter']),
                  'crmask diffs', len(d), d[:6].tolist(), 'clean diffs', len(dc), dc[:6].tolist(), 'status', ig.get('lazy_status'))
            for y, x in dc[:3]:
                print('   clean', y, x, 'gpu', clean_g[y, x], 'oracle', clean_o[y, x], 'in', img[y, x], 'cr', cr_o[y, x],
                      'nb cr', int(cr_o[max(y - 2, 0):y + 3, max(x - 2, 0):x + 3].sum()), 'bg', io.get('background'))


def chain():
    tel = 'BG3'
    raw = synth.make_raw(tel, 4001)[0]
    red = (2 * set_bb.ysize_chan, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, 9, red)
    coeffs = synth.make_xtalk(3)[3]
    raw_t = bbr._to_dev(raw)
    outs = []
    pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=4)
    for rep in range(3):
        snap = {}
        img = torch.empty(red, dtype=torch.float32, device='cuda')
        mask = torch.empty(red, dtype=torch.uint8, device='cuda')
        pipe._overscan(raw_t)
        torch.cuda.synchronize()
        snap['st'] = pipe.st.buf.clone()
        snap['means'] = pipe.means.clone()
        tel_, geom = pipe.tel, pipe.geom
        bbr.apply_enqueue(raw_t, geom, tel_, st=pipe.st, gain=pipe.gain, mbias=pipe.mbias, mflat=pipe.mflat, bpm=pipe.bpm,
                          want_mask=True, out_img=img, out_mask=mask, mwork=pipe.mwork)
        torch.cuda.synchronize()
        snap['apply_img'], snap['apply_mask'] = img.clone(), mask.clone()
        bbr.mask_morph_enqueue(mask, tel_, pipe.mwork)
        torch.cuda.synchronize()
        snap['morph_mask'] = mask.clone()
        for niter in (3, 4):
            im2 = img.clone()
            bbr.lacosmic_enqueue(im2, mask, pipe.crmask, 20, 0.01, 3, 0.0, niter, pipe.lwork, readnoise_dev=pipe.means[1:])
            torch.cuda.synchronize()
            snap['lac%d_img' % niter], snap['lac%d_cr' % niter] = im2.clone(), pipe.crmask.clone()
            snap['lac%d_info' % niter] = pipe.lwork.info.clone()
        outs.append(snap)
    for k in outs[0]:
        for rep in (1, 2):
            a, b = outs[0][k], outs[rep][k]
            same = torch.equal(a, b) if not a.is_floating_point() else bool(((a == b) | (a.isnan() & b.isnan())).all())
            extra = ''
            if not same:
                nd = int((a != b).sum())
                extra = ' ndiff %d' % nd
                if a.dim() == 2:
                    idx = torch.nonzero(a != b)[:4].tolist()
                    extra += ' at %s: %s vs %s' % (idx, [a[i[0], i[1]].item() for i in idx], [b[i[0], i[1]].item() for i in idx])
            print(k, 'rep', rep, 'same' if same else 'DIFF' + extra)
    print('info', outs[0]['lac4_info'].tolist(), outs[1]['lac4_info'].tolist())
    print('st std_vos', pipe.st.std_vos.tolist())


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'both'
    if which in ('stress', 'both'):
        stress()
    if which in ('chain', 'both'):
        chain()
