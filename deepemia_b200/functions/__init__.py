"""Host-side mirrors of the reference's src/functions modules on the hot path (same function names and signatures)."""
