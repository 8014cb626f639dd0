"""CPU oracle for the RIDERS radar-to-dense-depth hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``riders_b200/`` may import this
package; the only legal callers are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Each module restates one row of SURVEY.md section 8(a) in NumPy / CPU torch
and cites the reference file:line it follows.  The reference (MMOCKING/RIDERS)
is pure Python and ships no tests or golden vectors, so the oracle is pinned
against *outputs of the reference itself executed in the build container*:
``oracle/gen_golden.py`` imports ``/root/reference`` (when present), runs the
unmodified reference functions on seeded synthetic inputs and stores small
input/output fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py``
then checks every oracle function against those fixtures on any box.

Third-party arithmetic the reference delegates to (and which IS installed in
this image, so the oracle may call it on the CPU):
  * torchvision.ops.roi_pool   (reference pins torchvision 0.14; image has 0.26)
  * scipy.interpolate.griddata (reference pins scipy 1.10.1;  image has 1.18)
  * scipy.optimize.minimize_scalar
"""
