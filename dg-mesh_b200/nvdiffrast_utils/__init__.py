"""Host-side mirror of the reference's `nvdiffrast_utils` package for the hot path (DPSR)."""
