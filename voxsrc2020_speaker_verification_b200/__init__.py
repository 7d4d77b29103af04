"""B200-native inference hot path of xx205/voxsrc2020_speaker_verification: embedding extraction
(TDNN / Res2Net / DPN) and cosine + AS-norm scoring behind the reference's own stage interfaces."""
from . import arch  # noqa: F401

__all__ = ["arch", "Extractor", "Scorer"]


def __getattr__(name):
    if name == "Extractor":
        from .extractor import Extractor
        return Extractor
    if name == "Scorer":
        from .scoring import Scorer
        return Scorer
    raise AttributeError(name)
