"""fscl_b200: B200-native CLR sweep scan + block-permutation test (hot path of slowkoni/fscl)."""
