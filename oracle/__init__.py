"""CPU oracle for the tiled-detection hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference algorithm
(Abolfazlmsl/Oriented-Object-Detection, ``Detect_OBB.py`` / ``Train_OBB.py``) used as the
checker for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
package ``oriented_object_detection_b200`` never imports from here and has no CPU
fallback: it raises if the CUDA library is missing.

Pinning status:
  * pixel stages (tile plan, 3-ch copy, DT-Edge channel): pinned against the reference's
    own functions, lifted from ``/root/reference/Detect_OBB.py`` with ``ast`` and run
    unmodified on ``Input/Test1.png`` / ``Test2.png`` (``oracle/lift_reference.py``;
    vectors committed under ``tests/golden/``).
  * strike angle, NMS fixed point, output order: pinned by the 44 rows of the reference's
    ``Output/Test{1,2}.xlsx``.
  * rotated IoU (shapely/GEOS 2.0.7) and the Ultralytics 8.3.196 decode: the arithmetic
    lives in third-party packages absent from ``/root/reference`` and not installable
    offline -> restated from their published algorithms; PARITY UNPINNED for those two
    (indirectly pinned only through the xlsx fixed-point property).
"""
