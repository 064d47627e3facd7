"""CPU oracle for the belacks/audio-denoising inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker (or as the timed CPU baseline), never as a compute path of
``audio_denoising_b200``.

What it is: a torch-CPU fp32 restatement of

* the DSP chain whose arithmetic lives in torchaudio 2.6.0 (pinned by the
  reference's ``requirements.txt:3-4``; the image carries torchaudio 2.11.0,
  which is the oracle of record): ``Spectrogram``, ``MelScale``,
  ``InverseMelScale``, ``GriffinLim``, ``InverseSpectrogram`` as constructed at
  ``app3.py:135-153`` / ``server.py:173-176``;
* the reference's own model ``gruunet2.py:54-306``;
* the glue of ``app3.py:167-226`` (streaming hop) and ``server.py:200-216``.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md §4,
§8c).  The oracle is therefore pinned against outputs of the reference itself,
run in the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference`` under I/O stubs) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays them, and ``tests/test_oracle_torchaudio.py``
checks every DSP restatement against the live torchaudio transforms.
"""

from . import dsp, model, pipeline, synth, metrics  # noqa: F401
