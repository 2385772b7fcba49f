"""CPU oracle for the CLAP2Diffusion hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``clap2diffusion_b200``) may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may call into it, and only as the checker or as the
timed CPU baseline -- never as the thing shipped.

Parity status
-------------
* Audio side (``oracle/audio.py``): restates ``models/audio_adapter_v4.py``,
  ``models/hierarchical_audio_v4.py`` and ``models/audio_attention_processor.py``
  of the reference.  PINNED: ``oracle/make_golden.py`` imports the unmodified
  reference modules in the build container, loads the same synthetic weights and
  stores input/output vectors under ``tests/golden/``; ``tests/test_oracle_golden.py``
  replays them.
* SD-1.5 UNet / DDIM / Euler / VAE decoder (``oracle/sd15.py``): restates
  third-party ``diffusers==0.23.1`` (requirements.txt:7 of the reference), which is
  NOT vendored under /root/reference and not installable here.  PARITY UNPINNED
  beyond the exact parameter counts (859,520,964 / 49,490,179+20) and the
  attention-site census; see DESIGN.md.
"""
