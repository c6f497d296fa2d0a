"""No-op stub (the image has no matplotlib; triton.testing.perf_report imports pyplot unconditionally)."""
