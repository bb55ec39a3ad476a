"""CPU ORACLE -- test infrastructure only.

A CPU restatement of the reference's numpy/astropy/astroscrappy reduction path
(SURVEY.md section 8c).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product package
``blackbox_b200`` never does.

Pinning status ("Oracle"):
  * everything the reference computes itself (define_sections, gain_corr, os_corr, mask_init,
    cosmics_corr's wrapper, mask_header, xtalk_corr, nonlin_corr, and master_prep with its file
    selection, stack median, flat post-fix and header statistics) is pinned against the REFERENCE'S OWN CODE
    executed with stub modules for its absent dependencies: tests/golden/make_reference_golden.py
    -> tests/golden/reference_golden.json -> tests/test_reference_golden.py, tests/test_masters.py
    (bit for bit).
  * astropy sigma clipping and astroscrappy.detect_cosmics are restated from their published
    algorithms; the libraries are absent here and the reference has no tests or golden
    vectors -> PARITY UNPINNED for those two pieces (sigma_clip with a mean centre is checked
    against scipy.stats.sigmaclip).
  * rice.py (the Rice coder of fpacked raw frames, CFITSIO's ricecomp.c behind astropy in the
    reference's read_hdulist) is restated from the published algorithm -> PARITY UNPINNED beyond
    the hand-derived vectors in tests/test_zz_rice_fz.py (no fpack / CFITSIO / astropy here).
"""
