DEFAULT_SEED = 42  # constants/metadata.py:1 in the reference
