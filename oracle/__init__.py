"""CPU oracle for the DRSA/LRP hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / the timed CPU baseline.
The product path (``cxai`` + ``drsa_audio_b200``) never imports this package and
raises if the CUDA library is missing.
"""
