"""Mirror of the reference's `decoding/` package (same module and function names)."""
