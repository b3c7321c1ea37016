"""CPU oracle for the DeepMerge region-merging hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
or as the CPU arm being timed.  ``deepmerge_b200`` never imports it.

Pinning status (see DESIGN.md "Oracle"):

* pinned by executing the reference itself (``oracle/gen_golden.py`` imports
  ``/root/reference`` in place and writes ``tests/golden/*.npz``):
  ``Euclidean_distance`` / ``MC_Lyu_2020`` (ExtractFeatures.py:119-147, :228-237),
  the gather + ``np.mean`` pooling loop (ExtractFeatures.py:188-216),
  ``Nets.MLP.forward`` (Nets.py:28-35), ``Losses.Loss.forward`` (Losses.py:34-38),
  the RAG edge-list reader (MyUtils2.py:155-193), the pair sampler
  (MyUtils1.py:236-295) and the window geometry (MyUtils2.py:300-383).
* PARITY UNPINNED (the reference holds no code, tests or fixtures for them;
  the restatement follows the written spec in SURVEY.md section 8(a)):
  raster -> RAG, area / perimeter / band statistics, the merge loop, the
  pair-MLP input layout.  For those the oracle is cross-checked by two
  independent implementations (numpy here, C in ``dm_oracle.c``, scipy
  connected components in the tests).
"""
