"""CPU ORACLE -- test infrastructure only.

A CPU restatement of the reference's numpy/astropy/astroscrappy reduction path
(SURVEY.md section 8c).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product package
``blackbox_b200`` never does.

Pinning status ("Oracle"):
  * numpy / scipy pieces (np.median, np.polyfit, np.matmul, scipy.ndimage morphology,
    UnivariateSpline) call the very libraries the reference calls -> pinned by construction.
  * astropy sigma clipping and astroscrappy.detect_cosmics are restated from their published
    algorithms; the libraries are absent here and the reference has no tests or golden
    vectors -> PARITY UNPINNED for those two pieces.
"""
