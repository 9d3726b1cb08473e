"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the OptNetAlign hot path.

Two checkers live here:

* ``oracle.port``   -- ctypes binding of ``libona_oracle.so`` (oracle/ona_oracle.c), the
  plain-C restatement of the reference's ``Alignment.cpp`` / ``localSearch.cpp`` hot path.
* ``oracle.ref``    -- ctypes binding of ``_ref/libona_ref.so``: the UNMODIFIED reference
  sources compiled from /root/reference with ``oracle/ref_shim.cpp`` (available where the
  prebuilt library travelled; ``ref.available()`` says so).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
``optnetaligncpp_b200`` never does.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False):
    """Compile the C restatement (always) and _ref (only when /root/reference exists)."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout + out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")
