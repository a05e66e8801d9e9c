"""Throw-away stand-in for the `diffusers` package, used ONLY by tests/golden/make_golden.py to import the
reference's unmodified scheduler files in the build container (diffusers itself is not installed here).
The base class is the restatement in oracle/ddim_base.py (diffusers==0.31.0, parity unpinned)."""
