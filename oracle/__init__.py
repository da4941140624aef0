"""TEST INFRASTRUCTURE ONLY — CPU oracle for the Tacotron2 decoder recurrence.

Nothing under ``genvox_b200/`` (the product) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / CPU baseline.

Contents
--------
philox.py          counter-based Philox4x32-10 dropout-mask provider (numpy), the
                   bit-exact host mirror of the in-kernel generator.
synth.py           deterministic synthetic weights / inputs (Philox-driven, so the
                   same tensors are rebuilt on any machine without torch's RNG).
decoder_oracle.py  torch-CPU (fp32 / fp64) restatement of the reference decoder
                   recurrence, every function citing the reference file:line.
ref_import.py      imports the unmodified reference from /root/reference (build
                   container only) with four stubbed third-party modules.
make_golden.py     runs the unmodified reference with replayable dropout masks and
                   writes tests/golden/*.npz  (the pin for the oracle).

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated here by
make_golden.py and committed under tests/golden/.
"""
