"""Mirror of the reference's `rework/` directory: rework/decoding.py."""
