"""CPU oracle for the EVCont FCI hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and there only as the checker / the timed CPU
baseline.  ``evcont_b200`` never imports this package; the product path fails
loudly when the CUDA library is missing.

Parity status
-------------
* ``oracle.subspace`` / ``oracle.gradients`` restate the reference's *numpy*
  code (``evcont/ab_initio_eigenvector_continuation.py``,
  ``evcont/electron_integral_utils.py``, ``evcont/ab_initio_gradients_loewdin.py``).
  They are pinned against the reference itself, imported from /root/reference
  under a stub ``pyscf`` module, by ``tests/golden/make_golden.py``; the resulting
  vectors are committed under ``tests/golden/``.
* ``oracle.cistring`` / ``oracle.trans_rdm`` restate PySCF's ``cistring`` and
  ``trans_rdm12`` (pyscf, PyPI, version unpinned by the reference's
  pyproject.toml:10; call site evcont/FCI_EVCont.py:121).  PySCF is not
  installed in this image and the reference holds no test for it, so this part
  is "parity unpinned" with respect to the PySCF binary; it is pinned instead
  against a brute-force second-quantisation evaluation of the published
  definitions, RDM sum rules, a Slater-Condon Hamiltonian identity and the
  hand-derived link-table golden rows of SURVEY.md Appendix A.2.
"""
