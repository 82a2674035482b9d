"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference sampler loop.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker / CPU baseline.
The product package ``mjhmc_b200`` never imports this package.
"""
