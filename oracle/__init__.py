"""TEST INFRASTRUCTURE ONLY — CPU fp32 oracle for the ConvLSTM encoder-forecaster hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and the CPU-baseline / ``--impl reference`` legs of ``bench.py`` may import it, and only
as the checker (never as the thing measured as ours, never as a fallback).
The shipped package ``satflow_b200`` never imports this package.
"""
