"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.
The product path (``disconet_b200``) never imports this package.
"""
