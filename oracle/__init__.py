"""CPU oracle for the feature-stack + KMeans hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement of the reference
algorithms (numpy / scikit-learn / OpenCV, plus a plain-C GLCM) used as the
*checker* by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing under
``rs_image_segmentation_b200/`` may import it: the product path is CUDA-only and
fails loudly when the extension is missing.

Parity status ("Oracle"):
  * a1-a6, a8, a9 (normalise, indices, PCA, KMeans): PINNED - the restatement is
    checked against outputs of the reference's own functions, imported unmodified
    from /root/reference in the build container by ``tests/golden/make_golden.py``
    and committed as ``tests/golden/*.npz``.
  * a7 (GLCM): the arithmetic lives in scikit-image (``graycomatrix`` /
    ``graycoprops``), a third-party dependency that is absent from
    /root/reference, unpinned in its requirements.txt and not installable here.
    PARITY UNPINNED by the reference itself; pinned by the scikit-image
    docstring known-answer example and by the reference's call site
    (modules/features/indices.py:264-316) only.
"""
