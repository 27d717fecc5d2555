"""CPU oracle for the native quadrotor step()/reset() hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``reinmav-gym_b200/``) imports, links or executes anything from here; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and only as the checker / the timed CPU arm.

Tiers (see DESIGN.md "Oracle"):

* T0-ref  ``oracle/refshim.py`` — executes the reference's own source text from
  ``/root/reference`` under three stub modules.  Exists only in the build
  container; used to generate ``tests/golden/*.npz`` (``oracle/make_golden.py``).
* T0      ``oracle/batch.py`` (``dtype=float64``) — vectorised float64 NumPy restatement
  of the four ``step()`` functions and ``control()`` (reference file:line cited per
  function); ``oracle/reinmav64.py`` for reinmav-v0.  Pinned to T0-ref by the committed
  golden fixtures.
* T0s     ``oracle/scalar.py`` — one-object-per-env float64 port with the
  reference's execution model; the CPU arm of ``bench.py``.
* T1      ``oracle/batch.py`` (``dtype=float32``) + ``oracle/sincos32.py`` — float32
  restatement with the kernel's operation
  order (explicit fused multiply-adds emulated exactly, oracle/fma32.py); the CUDA path must
  equal it bit for bit.
* Philox  ``oracle/philox.py`` — Philox4x32-10, pinned to the Random123 KATs.

Parity status: the reference's own tests assert nothing, so the
pin is "outputs of the reference itself run here" — the golden fixtures — plus
the hand-checked KATs of SURVEY.md Appendix D.
"""
