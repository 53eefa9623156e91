"""deepemia_b200 — B200-native (sm_100a) post-head hot path of deepEMIA: paste/threshold, mask-IoU de-dup,
spatial-constraint filtering and morphometry, behind a C ABI (libemia.so)."""
__version__ = "0.1.0"
