"""CLI drop-ins for autoregressive/sample/sample_c2i.py, sample_t2i.py and sample_c2i_ddp.py (same flags)."""
