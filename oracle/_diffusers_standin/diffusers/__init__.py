"""Minimal stand-in for the `diffusers` package so the UNMODIFIED reference file
models/audio_attention_processor.py (which does `from diffusers.models.attention_processor import
Attention, AttnProcessor`, line 10) can be imported by oracle/make_golden.py.  ORACLE-ONLY test
infrastructure; restates diffusers 0.23.1 `Attention` from memory."""
