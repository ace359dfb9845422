"""Input side of the training loop.  The reference's datasets (datasets/{kinetics,audioset}.py on top of video_db.py) decode
video / audio with PyAV and librosa, which are outside this package's scope; what the hot path needs is the
batch contract of video_db.py:219-265, provided here by a synthetic dataset of the same shapes."""
from .synthetic import SyntheticAV   # noqa: F401
