"""CPU oracle for the gnngls inference hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker or as the timed CPU
baseline.  The product package (``gnngls_b200``) never imports from here and
fails loudly when its CUDA library is missing.

Contents
--------
``model_port.py``  torch restatement of DGL-0.6.1 ``GATConv`` (third-party; not
                   vendored in the reference; pinned at Pipfile.lock:316-329)
                   and of ``gnngls/models.py:5-70``.
``gls_port.c``     plain-C restatement of ``gnngls/operators.py``,
                   ``gnngls/algorithms.py:9-18,111-195`` and
                   ``gnngls/__init__.py:17-21`` (built by ``oracle/Makefile``).
``gls_port.py``    ctypes binding for ``gls_port.c``.
``ref_shim.py``    imports the UNMODIFIED reference from ``/root/reference``
                   (build container only; does not travel to the GPU box).
``make_golden.py`` runs the unmodified reference through ``ref_shim`` and
                   writes ``tests/golden/*.npz``.

Parity status
-------------
Search half (operators / local_search / guided_local_search / nearest_neighbor
/ tour_cost): PINNED — ``gls_port.c`` is checked bit-for-bit against golden
vectors produced by the reference's own Python (tests/test_oracle_golden.py).

Model half: the reference's own ``models.py`` is executed unmodified for the
golden vectors, but its ``dgl.nn.GATConv`` dependency is absent (not
installable offline), so GATConv itself is a restatement of the published
DGL 0.6.1 algorithm: "parity unpinned" for that one third-party op.  Two
independent restatements (edge-list scatter and dense masked softmax) are
checked against each other.
"""
