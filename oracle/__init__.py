"""CPU oracle for the VidSeg per-clip hot path -- TEST INFRASTRUCTURE ONLY.

Plain numpy / torch-CPU restatements of the reference algorithms, each function
citing the reference file:line (or the pinned third-party source) it follows.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package; the product package
``vidseg_diffusion_b200`` never does (tests/test_boundary.py checks that).

Parity status: the reference ships no golden vectors for this path (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference code itself,
generated in the authoring container by ``tests/golden/make_goldens.py`` (which
imports /root/reference through ``oracle/ref_import.py``) and against the
installed scikit-learn 1.9.0 ``KMeans`` (the reference pins 1.5.0; the Lloyd /
k-means++ code paths restated here are unchanged between the two).
"""
