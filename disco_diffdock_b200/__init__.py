"""B200-native reverse-diffusion docking sampler: drop-in for the hot path of gcorso/disco-diffdock."""
