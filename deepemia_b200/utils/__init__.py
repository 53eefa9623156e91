"""Host-side mirrors of the reference's src/utils modules on the hot path (same function names and signatures)."""
