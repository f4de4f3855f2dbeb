"""CPU/eager restatement of the MDN_SfM loss path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (its ``cpu_baseline`` leg and the
``--impl reference`` arm) may import it, and only as the checker / baseline.
The product package ``mdn_sfm_b200`` never imports this package and has no CPU
path of its own.

Parity status: the reference ships no golden vectors or tests (SURVEY.md
section 4), so the restatement is pinned against the reference ITSELF, imported
in place from /root/reference by ``oracle/ref_loader.py`` in the build
container (``tests/test_oracle_vs_reference.py``, bit-exact on CPU), and through
the fixtures that ``oracle/make_golden.py`` wrote to ``tests/golden/`` from
the reference's own functions.
"""
