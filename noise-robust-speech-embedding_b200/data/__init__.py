from .augment import add_noise_to_speech  # noqa: F401
from .noisy_speech_dataset import (  # noqa: F401
    GpuBatchMixer,
    MixedBatchLoader,
    NoiseRobustSpeechDataset,
    TensorPairDataset,
    create_dataloaders,
)
from .prefetch import DevicePrefetcher  # noqa: F401
