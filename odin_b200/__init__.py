"""odin_b200 -- B200-native (sm_100a) speech front-end + GMM-UBM Baum-Welch
behind the API of trungnt13/odin-ai's ``odin.preprocessing`` / ``odin.ml.GMM``.

    from odin_b200 import preprocessing as pp, ml

See DESIGN.md for the path and its boundary, include/odin_b200.h for the C-ABI.
"""
__version__ = "0.1.0"

from . import preprocessing  # noqa: F401,E402
from . import ml  # noqa: F401,E402
