"""nrse_b200: B200-native BYOL noisy-view hot path (SNR mix + normalisation, WavLM conv feature encoder,
BYOL cosine loss, EMA target update) behind the module surface of sunYtokki/Noise-Robust-Speech-Embedding.

Importable as ``nrse_b200`` (the directory carries the repository's hyphenated name; the top-level
``nrse_b200/`` shim aliases it).  Submodules:

* ``nrse_b200.ops``      torch custom ops over the C-ABI library (csrc/libnrse_b200.so, include/nrse_b200.h)
* ``nrse_b200.data``     ``add_noise_to_speech``, ``NoiseRobustSpeechDataset``, ``create_dataloaders``, GPU batch mixer
* ``nrse_b200.models``   ``WavLMEncoder``, ``BYOLSpeechModel``, ``byol_loss``, ``ProjectionHead``, ``PredictionHead``
* ``nrse_b200.train``    the BYOL step (``train_one_epoch``), data-parallel wrapper
"""
__version__ = "0.1.0"
