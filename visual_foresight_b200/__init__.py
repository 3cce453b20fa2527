"""visual_foresight_b200 — B200-native visual-MPC planning engine (CEM loop + CDNA/conv-LSTM predictor).

Only the hot path named by BASELINE.json's north_star lives here: csrc/ (CUDA kernels + the C-ABI
libvfengine.so) and the host-side mirror of the reference's Policy / predictor plugin surface."""
__all__ = ["spec"]
