from .datasets import VideoDataSet  # noqa: F401
