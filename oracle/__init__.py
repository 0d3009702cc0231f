"""CPU oracle for the DB_text_minimal hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``db_text_minimal_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there
only as the checker or as the CPU baseline being timed, never as the shipped path.

Parity status: PINNED against outputs of the reference itself.  The reference
(``/root/reference``, pure Python/PyTorch) has no tests or golden vectors of its own
(SURVEY.md section 4), so ``oracle/make_golden.py`` imports the reference code
unmodified in the build container, runs it on seeded inputs and commits the
results under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every
function of this restatement against those fixtures.  One stage stays unpinned:
``SegDetectorRepresenter.unclip`` (pyclipper 1.1.0.post3 / Clipper 6.4.2 and
Shapely 1.7.0 are third-party, absent from the reference tree and from this image).
"""
