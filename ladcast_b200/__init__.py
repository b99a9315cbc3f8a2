"""ladcast_b200 — B200-native (sm_100a) implementation of LaDCast's ensemble autoregressive latent-diffusion
rollout, behind the reference's own Python call signatures (see DESIGN.md / INTEGRATION.md)."""
__version__ = "0.1.0"
