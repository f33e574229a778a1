"""B200-native batched-inference hot path for VariantFormer."""
