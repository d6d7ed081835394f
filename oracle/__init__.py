"""oracle/ — TEST INFRASTRUCTURE, NOT PRODUCT.

CPU restatements of the reference's SELD front-end (Zeudon/sound-event-localization-detection,
``dataset.py`` + ``smrl_seld_gaussian.py``) used only as the parity checker.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  The product package
(``sound-event-localization-detection_b200``) never imports it and has no CPU fallback.

Parity status (see DESIGN.md "Oracle"):

* log-mel (dataset.py:27-58), point labels (dataset.py:60-119), region labels
  (smrl_seld_gaussian.py:397-534), windowing (dataset.py:267-330), polar_to_grid (utils.py:77-90):
  PINNED — checked against outputs of the reference itself, generated in the build container by
  ``tests/golden/make_golden.py`` (which imports the reference from a scratch copy) and committed as
  ``tests/golden/*.npz``; plus the notebook known-answers listed in SURVEY.md §8(c).
* FOA intensity vectors, GCC-PHAT, normalisation scaler: the reference has no such code
   — **parity unpinned**; the restatements here follow the formulas in SURVEY.md
  §8(a) rows A7-A9 and are cross-checked only through physical invariants.
"""
