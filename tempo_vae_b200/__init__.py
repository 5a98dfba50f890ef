"""tempo_vae_b200 — B200-native (sm_100a) engine for the TEMPO-VAE train / encode hot path."""
