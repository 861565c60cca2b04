"""Mirror of ``odin.preprocessing`` (odin/preprocessing/__init__.py:1-4) for the
accelerated path: ``pp.make_pipeline``, ``pp.base.*``, ``pp.speech.*``,
``pp.FeatureProcessor``."""
from . import base, signal, speech  # noqa: F401
from .base import *  # noqa: F401,F403
from .base import make_pipeline, Pipeline, Extractor, ExtractorSignal  # noqa: F401
from .speech import *  # noqa: F401,F403
from .processor import FeatureProcessor  # noqa: F401
