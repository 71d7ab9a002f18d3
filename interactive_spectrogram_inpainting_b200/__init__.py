"""B200-native VQ-VAE-2 code-extraction hot path."""
