"""Caller-side (MFEM/Palace stand-in) mesh, space and coefficient builders for the synthetic configs."""
