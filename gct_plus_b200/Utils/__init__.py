from .dataset import DataloaderPreparation, DeviceDataLoader, TokenisedCorpus  # noqa: F401
